// Host-side runtime of the radar-cube encoder: one C call runs Encoder.forward
// (model/models_radar_encoder.py:216-241) as a fixed sequence of kernel launches on one stream.
//
// Data layout: activations are channels-last [frames, D*H*W voxels, C]. The residual stream stays fp32
// (buffers x / y), every GroupNorm(+swish) writes the bf16 operand (xb) that the following tcgen05 implicit-GEMM
// convolution reads through TMA, and each convolution writes fp32 with bias and residual fused in its epilogue:
//
//   ResnetBlock (:82-100)   x -> [stats, GN+swish] -> conv1 -> [stats, GN+swish] -> conv2 (+ x or nin_shortcut(x))
//   Downsample  (:37-41)    cast -> stride-2 conv (TMA zero-fill = the high-side pad)
//   AttnBlock   (:112-135)  [stats, GN] -> fused q|k|v 1x1 conv (GEMM) -> 64-voxel attention -> proj_out GEMM (+ x)
//
// The encoder depends only on the radar cube, so it runs once per sample instead of once per network
// evaluation as in the reference (EDMPrecond.forward calls it 35 times per sample).
#include <cuda_bf16.h>

#include "../../include/rald_b200.h"
#include "host.cuh"
#include "kernels.h"

namespace rald {

struct EncCtx {
  const rald_enc_weights* w;
  float* x;
  float* y;
  float* t;
  void* xb;
  double* stats;
  int nf;
  cudaStream_t st;
};

static int conv3(const EncCtx& c, const rald_enc_conv& cv, const void* in_bf16, const float* resid, float* out, int d,
                 int h, int wd, int stride) {
  return conv3d_cl(in_bf16, cv.w, cv.w_rows, cv.b, resid, out, c.nf, d, h, wd, cv.cin, cv.cout, stride, c.st);
}

static int norm_act(const EncCtx& c, const float* src, const rald_enc_norm& n, int64_t V, int C, int mode) {
  RALD_TRY(gn_stats(src, c.nf, V, C, c.w->groups, c.stats, c.st));
  return gn_apply(src, c.stats, n.g, n.b, c.xb, c.nf, V, C, c.w->groups, c.w->eps, mode, c.st);
}

static int resblock(EncCtx& c, const rald_enc_resblock& rb, int d, int h, int wd) {
  const int64_t V = (int64_t)d * h * wd;
  const int cin = rb.c1.cin, cout = rb.c1.cout;
  const float* shortcut = c.x;
  if (rb.nin.w != nullptr) {
    // nin_shortcut: 1x1x1 conv of the raw input = GEMM over voxels (:96-99)
    RALD_TRY(gn_apply(c.x, nullptr, nullptr, nullptr, c.xb, c.nf, V, cin, 0, 0.f, 2, c.st));
    RALD_TRY(gemm_bf16(c.xb, cin, rb.nin.w, cin, c.y, cout, rb.nin.b, nullptr, 0, (int)(c.nf * V), cout, cin, 1, 0,
                       c.st));
    shortcut = c.y;
  }
  RALD_TRY(norm_act(c, c.x, rb.n1, V, cin, 0));
  RALD_TRY(conv3(c, rb.c1, c.xb, nullptr, c.t, d, h, wd, 1));
  RALD_TRY(norm_act(c, c.t, rb.n2, V, cout, 0));
  if (rb.nin.w != nullptr) {
    RALD_TRY(conv3(c, rb.c2, c.xb, shortcut, c.y, d, h, wd, 1));
    float* tmp = c.x; c.x = c.y; c.y = tmp;
  } else {
    RALD_TRY(conv3(c, rb.c2, c.xb, shortcut, c.x, d, h, wd, 1));
  }
  return 0;
}

static int attnblock(EncCtx& c, const rald_enc_attnblock& ab, int d, int h, int wd) {
  const int n = d * h * wd;
  const int C = ab.proj.cout;
  RALD_REQUIRE(n <= 64, "radar encoder: AttnBlock over %d voxels (the kernel handles <= 64 = 8x4x2)", n);
  RALD_TRY(norm_act(c, c.x, ab.n, n, C, 1));
  RALD_TRY(gemm_bf16(c.xb, C, ab.qkv.w, C, c.t, 3 * C, ab.qkv.b, nullptr, 0, c.nf * n, 3 * C, C, 1, 0, c.st));
  RALD_TRY(enc_attn(c.t, c.xb, c.nf, n, C, c.st));
  RALD_TRY(gemm_bf16(c.xb, C, ab.proj.w, C, c.x, C, ab.proj.b, c.x, C, c.nf * n, C, C, 1, 0, c.st));
  return 0;
}

}  // namespace rald

using namespace rald;

extern "C" int rald_radar_encoder(const rald_enc_weights* w, const rald_enc_workspace* ws, const float* x, float* out,
                                  int B, int D, int H, int W, void* stream) {
  RALD_REQUIRE(w != nullptr && ws != nullptr, "radar_encoder: null weights/workspace");
  RALD_REQUIRE(w->n_levels >= 1 && w->n_levels <= RALD_ENC_MAX_LEVELS, "radar_encoder: %d levels", w->n_levels);
  RALD_REQUIRE(w->ch % 64 == 0, "radar_encoder: base width %d must be a multiple of 64 (tcgen05 K tile)", w->ch);
  RALD_REQUIRE(B > 0 && ws->max_frames > 0, "radar_encoder: frames=%d micro-batch=%d", B, ws->max_frames);
  const int shrink = 1 << (w->n_levels - 1);
  RALD_REQUIRE(D % shrink == 0 && H % shrink == 0 && W % shrink == 0,
               "radar_encoder: resolution %dx%dx%d is not divisible by %d", D, H, W, shrink);
  // widest activation: level l has V0 / 8^l voxels and at most ch * 2^l... channels; check every level
  const int64_t V0 = (int64_t)D * H * W;
  {
    int64_t need = V0 * w->ch;
    int64_t V = V0;
    for (int l = 0; l < w->n_levels; ++l) {
      for (int j = 0; j < w->level[l].n_blocks; ++j) {
        const int64_t e = V * w->level[l].block[j].c1.cout;
        if (e > need) need = e;
      }
      V /= 8;
    }
    const int64_t vlast = V0 / ((int64_t)shrink * shrink * shrink);
    if (vlast * 3 * w->mid1.c1.cout > need) need = vlast * 3 * w->mid1.c1.cout;
    RALD_REQUIRE(ws->elems >= need * ws->max_frames, "radar_encoder: workspace holds %lld elements, %lld needed",
                 (long long)ws->elems, (long long)(need * ws->max_frames));
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t out_per_frame = (V0 / ((int64_t)shrink * shrink * shrink)) * w->z_ch;
  for (int f0 = 0; f0 < B; f0 += ws->max_frames) {
    EncCtx c;
    c.w = w; c.x = ws->x; c.y = ws->y; c.t = ws->t; c.xb = ws->xb; c.stats = ws->stats; c.st = st;
    c.nf = (B - f0) < ws->max_frames ? (B - f0) : ws->max_frames;
    int d = D, h = H, wd = W;
    RALD_TRY(enc_conv_in(x + (int64_t)f0 * V0 * w->in_ch, w->conv_in_w, w->conv_in_b, c.x, c.nf, d, h, wd, w->in_ch,
                         w->ch, st));
    for (int l = 0; l < w->n_levels; ++l) {
      const rald_enc_level& lv = w->level[l];
      for (int j = 0; j < lv.n_blocks; ++j) {
        RALD_TRY(resblock(c, lv.block[j], d, h, wd));
        if (j < lv.n_attn) RALD_TRY(attnblock(c, lv.attn[j], d, h, wd));
      }
      if (lv.down.w != nullptr) {
        const int64_t V = (int64_t)d * h * wd;
        RALD_TRY(gn_apply(c.x, nullptr, nullptr, nullptr, c.xb, c.nf, V, lv.down.cin, 0, 0.f, 2, st));
        RALD_TRY(conv3(c, lv.down, c.xb, nullptr, c.y, d, h, wd, 2));
        float* tmp = c.x; c.x = c.y; c.y = tmp;
        d /= 2; h /= 2; wd /= 2;
      }
    }
    RALD_TRY(resblock(c, w->mid1, d, h, wd));
    RALD_TRY(attnblock(c, w->mid_attn, d, h, wd));
    RALD_TRY(resblock(c, w->mid2, d, h, wd));
    RALD_TRY(norm_act(c, c.x, w->norm_out, (int64_t)d * h * wd, w->conv_out.cin, 0));
    RALD_TRY(conv3(c, w->conv_out, c.xb, nullptr, out + (int64_t)f0 * out_per_frame, d, h, wd, 1));
  }
  return 0;
}

extern "C" int rald_conv3d_cl(const void* x_bf16, const void* w_packed, int w_rows, const float* bias,
                              const float* resid, float* out, int B, int D, int H, int W, int Cin, int Cout, int stride,
                              void* stream) {
  return conv3d_cl(x_bf16, w_packed, w_rows, bias, resid, out, B, D, H, W, Cin, Cout, stride,
                   static_cast<cudaStream_t>(stream));
}
extern "C" int rald_enc_conv_in(const float* x, const float* w, const float* bias, float* out, int B, int D, int H,
                                int W, int Cin, int Cout, void* stream) {
  return enc_conv_in(x, w, bias, out, B, D, H, W, Cin, Cout, static_cast<cudaStream_t>(stream));
}
extern "C" int rald_gn_stats(const float* x, int B, int64_t V, int C, int groups, double* stats, void* stream) {
  return gn_stats(x, B, V, C, groups, stats, static_cast<cudaStream_t>(stream));
}
extern "C" int rald_gn_apply(const float* x, const double* stats, const float* gamma, const float* beta,
                             void* out_bf16, int B, int64_t V, int C, int groups, float eps, int mode, void* stream) {
  return gn_apply(x, stats, gamma, beta, out_bf16, B, V, C, groups, eps, mode, static_cast<cudaStream_t>(stream));
}
extern "C" int rald_enc_attn(const float* qkv, void* out_bf16, int B, int n, int C, void* stream) {
  return enc_attn(qkv, out_bf16, B, n, C, static_cast<cudaStream_t>(stream));
}
