// VecSet autoencoder runtime pieces (model/models_ae.py):
//   rald_ae_stack      proj + depth x [x += attn8h(LN x); x += FF_geglu(LN x)]           decode :410-414
//   linear_smallk      fp32 Linear with K <= 64 inputs (proj: latent_dim -> dim)          :346, :410
//   ln_dot_rows        v'[row] = (LayerNorm(x[row]) * g + b) . w   (fp32)                 folded decoder value path
// The latent stack reuses the denoiser's tcgen05 GEMM / attention / LayerNorm kernels.
#include <cuda_bf16.h>
#include <math.h>

#include "../../include/rald_b200.h"
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

// out[T, 512] = x[T, K] Wt[K, 512] + b ; one warp per row, Wt staged in smem.
__global__ void __launch_bounds__(256)
linear_smallk_kernel(const float* __restrict__ x, int K, const float* __restrict__ wt, const float* __restrict__ b,
                     float* __restrict__ out, int64_t T) {
  extern __shared__ float s_w[];  // [K][512]
  for (int i = threadIdx.x; i < K * 128; i += blockDim.x)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
    bias[j] = b ? __ldg(reinterpret_cast<const float4*>(b) + j * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < T; row += (int64_t)gridDim.x * 8) {
    const float x0 = lane < K ? x[row * K + lane] : 0.f;
    const float x1 = lane + 32 < K ? x[row * K + lane + 32] : 0.f;
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = bias[j];
    for (int c = 0; c < K; ++c) {
      const float xc = __shfl_sync(0xffffffffu, c < 32 ? x0 : x1, c & 31);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w = reinterpret_cast<const float4*>(s_w + c * 512)[j * 32 + lane];
        acc[j].x = fmaf(xc, w.x, acc[j].x);
        acc[j].y = fmaf(xc, w.y, acc[j].y);
        acc[j].z = fmaf(xc, w.z, acc[j].z);
        acc[j].w = fmaf(xc, w.w, acc[j].w);
      }
    }
    float4* o = reinterpret_cast<float4*>(out + row * 512);
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j * 32 + lane] = acc[j];
  }
}

int linear_smallk(const float* x, int K, const float* wt, const float* b, float* out, int64_t T, int N,
                  cudaStream_t stream) {
  RALD_REQUIRE(N == 512, "linear_smallk: N=%d unsupported (512 only)", N);
  RALD_REQUIRE(K >= 1 && K <= 64, "linear_smallk: K=%d must be in [1, 64]", K);
  const int smem = K * 512 * sizeof(float);
  static int configured = 0;
  if (smem > configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(linear_smallk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  int64_t blocks = (T + 7) / 8;
  const int64_t cap = device_sm_count();
  if (blocks > cap) blocks = cap;
  linear_smallk_kernel<<<(unsigned)blocks, 256, smem, stream>>>(x, K, wt, b, out, T);
  RALD_LAUNCHED();
  return 0;
}

__global__ void __launch_bounds__(256)
ln_dot_rows_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                   const float* __restrict__ w, float* __restrict__ out, int64_t rows, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * 512);
  float4 v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = xr[j * 32 + lane];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.0f / 512);
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / 512) + eps);
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + j * 32 + lane);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + j * 32 + lane);
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w) + j * 32 + lane);
    acc += ((v[j].x * rstd * gg.x + bb.x) * ww.x + (v[j].y * rstd * gg.y + bb.y) * ww.y) +
           ((v[j].z * rstd * gg.z + bb.z) * ww.z + (v[j].w * rstd * gg.w + bb.w) * ww.w);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

int ln_dot_rows(const float* x, const float* g, const float* b, const float* w, float* out, int64_t rows, int D,
                float eps, cudaStream_t stream) {
  RALD_REQUIRE(D == 512, "ln_dot_rows: D=%d unsupported (512 only)", D);
  const int64_t blocks = (rows + 7) / 8;
  ln_dot_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, g, b, w, out, rows, eps);
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

using namespace rald;

extern "C" int rald_ae_stack(const rald_ae_weights* w, const rald_dit_workspace* ws, const float* z, float* x_out,
                             int frames, void* stream) {
  RALD_REQUIRE(w != nullptr && ws != nullptr, "ae_stack: null weights/workspace");
  RALD_REQUIRE(w->dim == 512 && w->heads * 64 == w->dim, "ae_stack: dim=%d heads=%d unsupported (512 = 8 x 64)",
               w->dim, w->heads);
  RALD_REQUIRE(w->n_latents % 128 == 0 && w->n_latents <= 512, "ae_stack: n_latents=%d unsupported", w->n_latents);
  RALD_REQUIRE(frames > 0 && ws->max_frames > 0, "ae_stack: frames=%d", frames);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dim = w->dim, M = w->n_latents, heads = w->heads;
  const float scale = 1.0f / sqrtf((float)(dim / heads));
  GemmStaticWeights static_w;  // every GEMM below multiplies activations with packed model weights
  for (int f0 = 0; f0 < frames; f0 += ws->max_frames) {
    const int nf = (frames - f0) < ws->max_frames ? (frames - f0) : ws->max_frames;
    const int64_t T = (int64_t)nf * M;
    float* h = x_out + (int64_t)f0 * M * dim;  // the residual stream lives directly in the output buffer
    if (w->proj_wt != nullptr) {
      RALD_TRY(linear_smallk(z + (int64_t)f0 * M * w->latent_dim, w->latent_dim, w->proj_wt, w->proj_b, h, T, dim, st));
    } else {
      // deterministic AutoEncoder.decode (models_ae.py:260-264): the latents ARE the residual stream, no projection
      RALD_REQUIRE(w->latent_dim == dim, "ae_stack: latents of width %d need a projection to %d", w->latent_dim, dim);
      RALD_CHECK_CUDA(cudaMemcpyAsync(h, z + (int64_t)f0 * M * dim, sizeof(float) * T * dim, cudaMemcpyDeviceToDevice, st));
    }
    const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(ws->qkv);
    // precise: split weight pairs [rows][2 cols] (twice the elements per layer) + erf GELU, see rald_ae_weights
    const int wm = w->precise ? 2 : 1;
    for (int n = 0; n < w->depth; ++n) {
      const __nv_bfloat16* w_qkv = reinterpret_cast<const __nv_bfloat16*>(w->w_qkv) + (int64_t)n * 3 * dim * dim * wm;
      const __nv_bfloat16* w_o = reinterpret_cast<const __nv_bfloat16*>(w->w_o) + (int64_t)n * dim * dim * wm;
      const __nv_bfloat16* w_ff1 = reinterpret_cast<const __nv_bfloat16*>(w->w_ff1) + (int64_t)n * 8 * dim * dim * wm;
      const __nv_bfloat16* w_ff2 = reinterpret_cast<const __nv_bfloat16*>(w->w_ff2) + (int64_t)n * 4 * dim * dim * wm;
      RALD_TRY(ln_rows(h, dim, w->ln1_w + (int64_t)n * dim, w->ln1_b + (int64_t)n * dim, 0, 0, 0, ws->xn, dim, 0, T,
                       dim, 1e-5f, st));
      if (w->precise) {
        RALD_TRY(gemm_bf16_wsplit(ws->xn, dim, w_qkv, 2 * dim, ws->qkv, 3 * dim, nullptr, nullptr, 0, (int)T, 3 * dim,
                                  dim, 0, 2 * dim, 3 * dim, 0, st));
      } else {
        RALD_TRY(gemm_bf16_f16cols(ws->xn, dim, w_qkv, dim, ws->qkv, 3 * dim, nullptr, (int)T, 3 * dim, dim, 2 * dim,
                                   3 * dim, st));
      }
      RALD_TRY(attn_d64(qkv, 3 * dim, qkv + dim, 3 * dim, qkv + 2 * dim, 3 * dim, ws->att, dim, nf, heads, M, M, scale,
                        st));
      if (w->precise) {
        RALD_TRY(gemm_bf16_wsplit(ws->att, dim, w_o, 2 * dim, h, dim, w->b_o + (int64_t)n * dim, h, dim, (int)T, dim,
                                  dim, 1, 0, 0, 0, st));
      } else {
        RALD_TRY(gemm_bf16(ws->att, dim, w_o, dim, h, dim, w->b_o + (int64_t)n * dim, h, dim, (int)T, dim, dim, 1, 0,
                           st));
      }
      RALD_TRY(ln_rows(h, dim, w->ln2_w + (int64_t)n * dim, w->ln2_b + (int64_t)n * dim, 0, 0, 0, ws->xn, dim, 0, T,
                       dim, 1e-5f, st));
      if (w->precise) {
        RALD_TRY(gemm_bf16_wsplit(ws->xn, dim, w_ff1, 2 * dim, ws->ff, 4 * dim, w->b_ff1 + (int64_t)n * 8 * dim, nullptr,
                                  0, (int)T, 8 * dim, dim, 2, 0, 0, 1, st));
        RALD_TRY(gemm_bf16_wsplit(ws->ff, 4 * dim, w_ff2, 8 * dim, h, dim, w->b_ff2 + (int64_t)n * dim, h, dim, (int)T,
                                  dim, 4 * dim, 1, 0, 0, 0, st));
      } else {
        RALD_TRY(gemm_bf16(ws->xn, dim, w_ff1, dim, ws->ff, 4 * dim, w->b_ff1 + (int64_t)n * 8 * dim, nullptr, 0, (int)T,
                           8 * dim, dim, 2, 0, st));
        RALD_TRY(gemm_bf16(ws->ff, 4 * dim, w_ff2, 4 * dim, h, dim, w->b_ff2 + (int64_t)n * dim, h, dim, (int)T, dim,
                           4 * dim, 1, 0, st));
      }
    }
  }
  return 0;
}

extern "C" int rald_linear_smallk(const float* x, int K, const float* wt, const float* b, float* out, int64_t T,
                                  int N, void* stream) {
  return linear_smallk(x, K, wt, b, out, T, N, static_cast<cudaStream_t>(stream));
}

extern "C" int rald_ln_dot_rows(const float* x, const float* g, const float* b, const float* w, float* out,
                                int64_t rows, int D, float eps, void* stream) {
  return ln_dot_rows(x, g, b, w, out, rows, D, eps, static_cast<cudaStream_t>(stream));
}

extern "C" int rald_ae_query(const float* queries, int B, int64_t Q, const void* wpe_bf16, const float* pe_bias,
                             const float* ln_g, const float* ln_b, const void* kprime_bf16, const float* vprime,
                             const float* c0, const float* freq24_host, float* logits, int dim, int n_latents,
                             void* stream) {
  return ae_query(queries, B, Q, wpe_bf16, pe_bias, ln_g, ln_b, kprime_bf16, vprime, c0, freq24_host, logits, dim,
                  n_latents, static_cast<cudaStream_t>(stream));
}
