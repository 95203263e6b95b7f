// Backward-pass kernels of the radar-cube encoder (training with `unfreeze_radar_enc: true`, SURVEY.md §8(f) row 3;
// reference: autograd through model/models_radar_encoder.py:5-12 (Normalize, swish), :29-44 (Downsample), :46-100
// (ResnetBlock), :102-135 (AttnBlock), :137-241 (Encoder)). All activations channels-last [B, V, C].
//
// The matrix work of the encoder's backward pass runs on the existing tensor-core kernels:
//   conv dgrad  = the forward implicit-GEMM convolution (conv3d.cu) on spatially flipped, in/out-transposed weights
//                 (stride 2: over the output gradient zero-stuffed onto the input grid, enc_stuff_kernel);
//   conv wgrad  = ONE split-K GEMM launch over K = all voxels of the batch: A = the three kw-shifted copies of X, W = dY,
//                 both written once, transposed, onto the zero-PADDED voxel grid (enc_pad_transpose_kernel), where a tap
//                 is a constant offset of the flattened index. The nine (kd, kh) taps are nine column blocks of the
//                 output, each reading dY^T at its own column offset (multiples of 8: TMA box origins must be 16-byte
//                 aligned, hence the pre-shifted kw copies); four taps (Cout = 64) share one 256-wide tile and one A tile
//                 (gemm.cu: tap_n / tap_shift; out-of-range coordinates read zeros);
//   1x1 convs   = GEMMs as in the denoiser (train_bwd.cu helpers).
// This file holds the rest: GroupNorm(+swish) backward in two passes, the padded / dilated transpose, the zero-stuffing
// and the backward of the 64-voxel single-head attention.
#include "../../include/rald_b200.h"

#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------------------
// GroupNorm (+ swish) backward. y = act(yh), yh = xh * gamma + beta, xh = (x - mean_g) * rstd_g.
// Pass 1: sums[b][c] = { sum_v dyh * xh, sum_v dyh } with dyh = dy * act'(yh)   (fp64 atomics across blocks)
// Pass 2: dx (+)= rstd_g * (gamma_c * dyh - m1_g - xh * m2_g) (+ add), m1_g = sum_{c in g} gamma_c sums[b][c][1] / n,
//         m2_g = sum_{c in g} gamma_c sums[b][c][0] / n, n = V * C / groups.
// d gamma_c = sum_b sums[b][c][0], d beta_c = sum_b sums[b][c][1] (summed by the caller: B x C numbers).
// ---------------------------------------------------------------------------------------------------
struct GnBwdParams {
  const float* x;
  const float* dy;
  const double* stats;   // [B][groups][2] = sum, sum of squares (forward pass)
  const float* gamma;
  const float* beta;
  double* sums;          // [B][C][2]
  const float* add;      // optional [B][V][C] added to dx (identity shortcut / residual gradient)
  float* dx;
  int64_t V;
  int C, groups;
  float eps;
  int swish;
  int64_t vox_per_block;
};

__device__ __forceinline__ void gn_channel_consts(const GnBwdParams& p, int b, float* s_mean, float* s_rstd) {
  const int cpg = p.C / p.groups;
  for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
    const double n = (double)p.V * cpg;
    const double mean = p.stats[((int64_t)b * p.groups + g) * 2 + 0] / n;
    double var = p.stats[((int64_t)b * p.groups + g) * 2 + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = (float)mean;
    s_rstd[g] = (float)(1.0 / sqrt(var + (double)p.eps));
  }
}

__device__ __forceinline__ float gn_dyh(float x, float dy, float mean, float rstd, float gamma, float beta, int swish,
                                        float& xh) {
  xh = (x - mean) * rstd;
  if (!swish) return dy;
  const float yh = fmaf(xh, gamma, beta);
  const float s = sigmoidf_(yh);
  return dy * (s * (1.0f + yh * (1.0f - s)));
}

__global__ void __launch_bounds__(256)
gn_bwd_reduce_kernel(const GnBwdParams p) {
  __shared__ float s_mean[256], s_rstd[256];
  __shared__ float s_a[256 * 4], s_b[256 * 4];
  const int b = blockIdx.y;
  gn_channel_consts(p, b, s_mean, s_rstd);
  __syncthreads();
  const int C = p.C, quads = C >> 2, cpg = C / p.groups;
  const int vlanes = blockDim.x / quads;
  const int cq = threadIdx.x % quads, vl = threadIdx.x / quads;
  const int64_t v0 = (int64_t)blockIdx.x * p.vox_per_block;
  const int64_t v1 = (v0 + p.vox_per_block) < p.V ? (v0 + p.vox_per_block) : p.V;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, bb[4] = {0.f, 0.f, 0.f, 0.f};
  if (vl < vlanes) {
    const float4* xb = reinterpret_cast<const float4*>(p.x + (int64_t)b * p.V * C);
    const float4* db = reinterpret_cast<const float4*>(p.dy + (int64_t)b * p.V * C);
    float gm[4], bt[4], mean[4], rstd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = cq * 4 + k;
      gm[k] = p.gamma[c]; bt[k] = p.beta[c]; mean[k] = s_mean[c / cpg]; rstd[k] = s_rstd[c / cpg];
    }
    for (int64_t v = v0 + vl; v < v1; v += vlanes) {
      const float4 xv = xb[v * quads + cq], dv = db[v * quads + cq];
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float xh;
        const float g = gn_dyh(xs[k], ds[k], mean[k], rstd[k], gm[k], bt[k], p.swish, xh);
        a[k] = fmaf(g, xh, a[k]);
        bb[k] += g;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      s_a[vl * C + cq * 4 + k] = a[k];
      s_b[vl * C + cq * 4 + k] = bb[k];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double sa = 0.0, sb = 0.0;
    for (int l = 0; l < vlanes; ++l) { sa += (double)s_a[l * C + c]; sb += (double)s_b[l * C + c]; }
    atomicAdd(&p.sums[((int64_t)b * C + c) * 2 + 0], sa);
    atomicAdd(&p.sums[((int64_t)b * C + c) * 2 + 1], sb);
  }
}

__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const GnBwdParams p) {
  __shared__ float s_mean[256], s_rstd[256], s_m1[256], s_m2[256];
  const int b = blockIdx.y;
  const int C = p.C, cpg = C / p.groups, quads = C >> 2;
  gn_channel_consts(p, b, s_mean, s_rstd);
  for (int g = threadIdx.x; g < p.groups; g += blockDim.x) {
    double m1 = 0.0, m2 = 0.0;
    for (int j = 0; j < cpg; ++j) {
      const int c = g * cpg + j;
      m2 += (double)p.gamma[c] * p.sums[((int64_t)b * C + c) * 2 + 0];
      m1 += (double)p.gamma[c] * p.sums[((int64_t)b * C + c) * 2 + 1];
    }
    const double n = (double)p.V * cpg;
    s_m1[g] = (float)(m1 / n);
    s_m2[g] = (float)(m2 / n);
  }
  __syncthreads();
  const int64_t nq = p.V * quads;
  const int64_t q0 = (int64_t)blockIdx.x * p.vox_per_block;   // (here: quads per block)
  const int64_t q1 = (q0 + p.vox_per_block) < nq ? (q0 + p.vox_per_block) : nq;
  const float4* xb = reinterpret_cast<const float4*>(p.x + (int64_t)b * p.V * C);
  const float4* db = reinterpret_cast<const float4*>(p.dy + (int64_t)b * p.V * C);
  const float4* ab = p.add ? reinterpret_cast<const float4*>(p.add + (int64_t)b * p.V * C) : nullptr;
  float4* ob = reinterpret_cast<float4*>(p.dx + (int64_t)b * p.V * C);
  for (int64_t i = q0 + threadIdx.x; i < q1; i += blockDim.x) {
    const int c0 = (int)(i % quads) * 4;
    const float4 xv = xb[i], dv = db[i];
    const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + k, g = c / cpg;
      float xh;
      const float gm = p.gamma[c];
      const float dyh = gn_dyh(xs[k], ds[k], s_mean[g], s_rstd[g], gm, p.beta[c], p.swish, xh);
      o[k] = s_rstd[g] * (gm * dyh - s_m1[g] - xh * s_m2[g]);
    }
    if (ab != nullptr) {
      const float4 av = ab[i];
      o[0] += av.x; o[1] += av.y; o[2] += av.z; o[3] += av.w;
    }
    ob[i] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int gn_bwd(const float* x, const float* dy, const double* stats, const float* gamma, const float* beta, int B, int64_t V,
           int C, int groups, float eps, int swish, double* sums, const float* add, float* dx, cudaStream_t stream) {
  RALD_REQUIRE(C % 4 == 0 && C <= 256 && C >= 4 && 256 % (C / 4) == 0, "gn_bwd: C=%d unsupported", C);
  RALD_REQUIRE(groups > 0 && groups <= 256 && C % groups == 0, "gn_bwd: groups=%d does not divide C=%d", groups, C);
  RALD_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * B * C, stream));
  GnBwdParams p;
  p.x = x; p.dy = dy; p.stats = stats; p.gamma = gamma; p.beta = beta; p.sums = sums; p.add = add; p.dx = dx;
  p.V = V; p.C = C; p.groups = groups; p.eps = eps; p.swish = swish;
  p.vox_per_block = (16384 * 4) / C;
  if (p.vox_per_block < 1) p.vox_per_block = 1;
  {
    dim3 grid((unsigned)((V + p.vox_per_block - 1) / p.vox_per_block), (unsigned)B);
    ProfScope prof(FAM_GN, stream, (double)B * V * C * 8.0);
    gn_bwd_reduce_kernel<<<grid, 256, 0, stream>>>(p);
    RALD_LAUNCHED();
  }
  {
    p.vox_per_block = 256 * 16;   // float4 quads per block
    const int64_t nq = V * (C / 4);
    dim3 grid((unsigned)((nq + p.vox_per_block - 1) / p.vox_per_block), (unsigned)B);
    ProfScope prof(FAM_GN, stream, (double)B * V * C * (add ? 16.0 : 12.0));
    gn_bwd_apply_kernel<<<grid, 256, 0, stream>>>(p);
    RALD_LAUNCHED();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// in [B, D, H, W, C] (fp32 or bf16, channels-last) -> out_t bf16 [copies * copy_rows][ld]: element (c, P) of copy r at
// row r * copy_rows + c, column P + pos_bias - r, with P the flattened index of voxel (dil*d + 1, dil*h + 1, dil*w + 1)
// on the zero-padded target grid [B][Dt+2][Ht+2][Wp] (Dt = dil * D, ...; row pitch Wp >= Wt + 2, a multiple of 8).
// `out_t` must be zero-initialised (pad positions and, for dil = 2, the stuffed zeros are never written).
// copies = 3, pos_bias = o: the three kw taps of a convolution as three pre-shifted copies — TMA needs 16-byte aligned
// box origins (a 2-byte shifted origin raises "illegal instruction", measured), so only multiples of 8 columns can be
// passed to the GEMM as an operand shift; with Wp % 8 == 0 the kd / kh taps are such multiples and kw is not.
// 64 voxels x 64 channels per CTA through shared memory.
// ---------------------------------------------------------------------------------------------------
template <bool IN_F32>
__global__ void __launch_bounds__(256)
enc_pad_transpose_kernel(const void* __restrict__ in, int64_t total_vox, int D, int H, int W, int C, int dil, int Wp,
                         int copies, int copy_rows, int64_t pos_bias, uint16_t* __restrict__ out_t, int64_t ld,
                         int panel_len, int halo, int n_panels, double* __restrict__ colsum) {
  __shared__ uint16_t tile[64][66];
  __shared__ int64_t s_p[64];
  __shared__ float s_cs[4][64];
  float cs = 0.f;
  const int64_t v0 = (int64_t)blockIdx.y * 64;
  const int c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  if (threadIdx.x < 64) {
    int64_t v = v0 + threadIdx.x;
    int64_t P = -1;
    if (v < total_vox) {
      const int w = (int)(v % W); v /= W;
      const int h = (int)(v % H); v /= H;
      const int d = (int)(v % D);
      const int64_t b = v / D;
      const int64_t Dp = (int64_t)dil * D + 2, Hp = (int64_t)dil * H + 2;
      P = ((b * Dp + (dil * d + 1)) * Hp + (dil * h + 1)) * Wp + (dil * w + 1);
    }
    s_p[threadIdx.x] = P;
  }
  // load: consecutive threads walk the channels of one voxel (contiguous in memory)
#pragma unroll 4
  for (int i = ty; i < 64; i += 4) {
    const int64_t v = v0 + i;
    const int c = c0 + tx;
    uint16_t val = 0;
    if (v < total_vox && c < C) {
      if (IN_F32) {
        const float f = reinterpret_cast<const float*>(in)[v * C + c];
        cs += f;
        __nv_bfloat16 hv = __float2bfloat16_rn(f);
        val = *reinterpret_cast<uint16_t*>(&hv);
      } else {
        val = reinterpret_cast<const uint16_t*>(in)[v * C + c];
      }
    }
    tile[i][tx] = val;
  }
  if (IN_F32 && colsum != nullptr) s_cs[ty][tx] = cs;
  __syncthreads();
  // bias gradient: column sums of the fp32 input, one fp64 atomic per (CTA, channel) (order-dependent at 1e-16 relative)
  if (IN_F32 && colsum != nullptr && threadIdx.x < 64 && c0 + threadIdx.x < C)
    atomicAdd(&colsum[c0 + threadIdx.x], (double)((s_cs[0][threadIdx.x] + s_cs[1][threadIdx.x]) +
                                                  (s_cs[2][threadIdx.x] + s_cs[3][threadIdx.x])));
  // store: consecutive threads walk the voxels of one channel row (contiguous runs of W on the padded grid)
#pragma unroll 4
  for (int i = ty; i < 64; i += 4) {
    const int c = c0 + i;
    const int64_t P = s_p[tx];
    if (c < C && P >= 0) {
      const uint16_t val = tile[tx][i];
      for (int r = 0; r < copies; ++r) {
        const int64_t pos = P + pos_bias - r;
        const int64_t orow = (int64_t)r * copy_rows + c;
        if (panel_len == 0) {
          out_t[orow * ld + pos] = val;
        } else {
          // K-panel-major: [panel][row][halo | panel_len | halo]; a column near a panel edge is repeated in the
          // neighbour's halo
          const int64_t rows_total = (int64_t)copies * copy_rows, pls = panel_len + 2 * halo;
          const int64_t pn = pos / panel_len;
          const int col = (int)(pos - pn * panel_len);
          out_t[(pn * rows_total + orow) * pls + col + halo] = val;
          if (col < halo && pn > 0) out_t[((pn - 1) * rows_total + orow) * pls + col + panel_len + halo] = val;
          if (col >= panel_len - halo && pn + 1 < n_panels)
            out_t[((pn + 1) * rows_total + orow) * pls + col - panel_len + halo] = val;
        }
      }
    }
  }
}

int enc_pad_transpose(const void* in, int in_f32, int B, int D, int H, int W, int C, int dil, int Wp, int copies,
                      int copy_rows, int pos_bias, void* out_t, int64_t ld, int panel_len, int halo, int n_panels,
                      double* colsum, cudaStream_t stream) {
  RALD_REQUIRE(colsum == nullptr || in_f32, "enc_pad_transpose: column sums are taken from an fp32 input");
  if (colsum != nullptr) RALD_CHECK_CUDA(cudaMemsetAsync(colsum, 0, sizeof(double) * C, stream));
  RALD_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && C > 0 && (dil == 1 || dil == 2), "enc_pad_transpose: bad geometry");
  RALD_REQUIRE(Wp >= dil * W + 2 && Wp % 8 == 0, "enc_pad_transpose: row pitch %d must be a multiple of 8 >= %d", Wp,
               dil * W + 2);
  RALD_REQUIRE((copies == 1 || copies == 3) && copy_rows >= C && pos_bias >= 0 && pos_bias <= 1,
               "enc_pad_transpose: bad copy layout");
  const int64_t padded = (int64_t)B * (dil * D + 2) * (dil * H + 2) * Wp;
  if (panel_len == 0) {
    RALD_REQUIRE(ld >= padded + 8, "enc_pad_transpose: row pitch %lld < padded grid %lld", (long long)ld, (long long)padded);
  } else {
    RALD_REQUIRE(panel_len % 64 == 0 && halo >= 0 && halo % 8 == 0 && halo < panel_len && n_panels > 0 &&
                 (int64_t)n_panels * panel_len >= padded + 8, "enc_pad_transpose: %d panels of %d columns (halo %d) do not "
                 "hold the padded grid of %lld", n_panels, panel_len, halo, (long long)padded);
  }
  // frame groups whose voxel-tile count fits gridDim.y
  const int64_t vox_per_frame = (int64_t)D * H * W;
  const int64_t max_frames = (65535ll * 64) / vox_per_frame;
  RALD_REQUIRE(max_frames >= 1, "enc_pad_transpose: one frame has too many voxels");
  const int esz = in_f32 ? 4 : 2;
  const int64_t padded_frame = padded / B;
  for (int64_t f0 = 0; f0 < B; f0 += max_frames) {
    const int64_t nf = (B - f0) < max_frames ? (B - f0) : max_frames;
    const int64_t tv = nf * vox_per_frame;
    dim3 g2((unsigned)((C + 63) / 64), (unsigned)((tv + 63) / 64));
    const char* src = reinterpret_cast<const char*>(in) + f0 * vox_per_frame * C * esz;
    const int64_t bias = pos_bias + f0 * padded_frame;     // the group's frames start at this column
    uint16_t* dst = reinterpret_cast<uint16_t*>(out_t);
    if (in_f32) enc_pad_transpose_kernel<true><<<g2, 256, 0, stream>>>(src, tv, D, H, W, C, dil, Wp, copies, copy_rows, bias,
                                                                       dst, ld, panel_len, halo, n_panels, colsum);
    else enc_pad_transpose_kernel<false><<<g2, 256, 0, stream>>>(src, tv, D, H, W, C, dil, Wp, copies, copy_rows, bias, dst,
                                                                 ld, panel_len, halo, n_panels, nullptr);
    RALD_LAUNCHED();
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Zero-stuffing for the stride-2 convolution's dgrad: in fp32 [B, D, H, W, C] -> out bf16 [B, 2D, 2H, 2W, C] with
// out[2d+1, 2h+1, 2w+1] = in[d, h, w]; `out` must be zero-initialised. One thread per (voxel, 4 channels).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
enc_stuff_kernel(const float4* __restrict__ in, int64_t total_quads, int D, int H, int W, int quads, uint2* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total_quads; i += (int64_t)gridDim.x * 256) {
    const int cq = (int)(i % quads);
    int64_t v = i / quads;
    const int w = (int)(v % W); v /= W;
    const int h = (int)(v % H); v /= H;
    const int d = (int)(v % D);
    const int64_t b = v / D;
    const int64_t o = ((b * (2 * D) + (2 * d + 1)) * (2 * H) + (2 * h + 1)) * (2 * W) + (2 * w + 1);
    const float4 x = in[i];
    out[o * quads + cq] = make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w));
  }
}

int enc_stuff(const float* in, int B, int D, int H, int W, int C, void* out_bf16, cudaStream_t stream) {
  RALD_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && C % 4 == 0, "enc_stuff: bad geometry");
  const int64_t total = (int64_t)B * D * H * W * (C / 4);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  enc_stuff_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(in), total, D, H, W, C / 4,
                                                         reinterpret_cast<uint2*>(out_bf16));
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Backward of enc_attn (AttnBlock core :121-133): qkv fp32 [B*n, 3C], dO fp32 [B*n, C] -> dqkv fp32 [B*n, 3C].
// One CTA per frame (n <= 64 voxels, one head of width C <= 256); probabilities and dS in shared memory, the operands
// are read from global memory (a frame's q, k, v, dO are 256 KB: L1 / L2 resident).
//   P = softmax(q k^T s), dP = dO v^T, dS = P o (dP - rowsum(P o dP)) s, dq = dS k, dk = dS^T q, dv = P^T dO.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
enc_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dO, float* __restrict__ dqkv, int n, int C,
                    float scale) {
  __shared__ float s_p[64 * 65];
  __shared__ float s_ds[64 * 65];
  const int b = blockIdx.x;
  const float* base = qkv + (int64_t)b * n * 3 * C;
  const float* dob = dO + (int64_t)b * n * C;
  float* outb = dqkv + (int64_t)b * n * 3 * C;
  const int quads = C >> 2;
  // (every CTA of a frame recomputes the 64 x 64 probabilities: 2 MFLOP; the CTAs split the channels of the last phase)
  // scores and dP: thread (i, jq) handles 16 columns of row i
  {
    const int i = threadIdx.x >> 2, jq = threadIdx.x & 3;
    for (int jj = 0; jj < 16; ++jj) {
      const int j = jq * 16 + jj;
      float acc = -INFINITY, dp = 0.f;
      if (i < n && j < n) {
        const float4* qa = reinterpret_cast<const float4*>(base + (int64_t)i * 3 * C);
        const float4* kb = reinterpret_cast<const float4*>(base + (int64_t)j * 3 * C + C);
        const float4* vb = reinterpret_cast<const float4*>(base + (int64_t)j * 3 * C + 2 * C);
        const float4* da = reinterpret_cast<const float4*>(dob + (int64_t)i * C);
        float a0 = 0.f, a1 = 0.f;
        for (int c = 0; c < quads; ++c) {
          const float4 q4 = qa[c], k4 = kb[c], v4 = vb[c], d4 = da[c];
          a0 = fmaf(q4.x, k4.x, a0); a0 = fmaf(q4.y, k4.y, a0); a0 = fmaf(q4.z, k4.z, a0); a0 = fmaf(q4.w, k4.w, a0);
          a1 = fmaf(d4.x, v4.x, a1); a1 = fmaf(d4.y, v4.y, a1); a1 = fmaf(d4.z, v4.z, a1); a1 = fmaf(d4.w, v4.w, a1);
        }
        acc = a0 * scale;
        dp = a1;
      }
      if (i < 64) { s_p[i * 65 + j] = acc; s_ds[i * 65 + j] = dp; }
    }
  }
  __syncthreads();
  // softmax rows and dS: 4 threads per row
  {
    const int i = threadIdx.x >> 2, jq = threadIdx.x & 3;
    const bool ok = i < n;
    float e[16], dpv[16];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      e[jj] = ok ? s_p[i * 65 + jq * 16 + jj] : -INFINITY;
      dpv[jj] = ok ? s_ds[i * 65 + jq * 16 + jj] : 0.f;
      mx = fmaxf(mx, e[jj]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      e[jj] = ok ? expf(e[jj] - mx) : 0.f;
      sum += e[jj];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = ok ? 1.0f / sum : 0.f;
    float dsum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      e[jj] *= inv;
      dsum = fmaf(e[jj], dpv[jj], dsum);
    }
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
    dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
    if (ok) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        s_p[i * 65 + jq * 16 + jj] = e[jj];
        s_ds[i * 65 + jq * 16 + jj] = e[jj] * (dpv[jj] - dsum) * scale;
      }
    }
  }
  __syncthreads();
  // dq[i] = sum_j dS[i][j] k[j];  dk[j] = sum_i dS[i][j] q[i];  dv[j] = sum_i P[i][j] dO[i]
  // thread <-> (channel of this CTA's slice, quarter of the rows)
  const int cpb = (C + gridDim.y - 1) / gridDim.y;           // channels per CTA
  const int rows_q = (n + 3) / 4;
  for (int cc = threadIdx.x & 63; cc < cpb; cc += 64) {
    const int c = blockIdx.y * cpb + cc;
    if (c >= C) break;
    const int i0 = (threadIdx.x >> 6) * rows_q;
    const int i1 = (i0 + rows_q) < n ? (i0 + rows_q) : n;
    for (int i = i0; i < i1; ++i) {
      float aq = 0.f, ak = 0.f, av = 0.f;
      for (int j = 0; j < n; ++j) {
        aq = fmaf(s_ds[i * 65 + j], base[(int64_t)j * 3 * C + C + c], aq);
        ak = fmaf(s_ds[j * 65 + i], base[(int64_t)j * 3 * C + c], ak);
        av = fmaf(s_p[j * 65 + i], dob[(int64_t)j * C + c], av);
      }
      outb[(int64_t)i * 3 * C + c] = aq;
      outb[(int64_t)i * 3 * C + C + c] = ak;
      outb[(int64_t)i * 3 * C + 2 * C + c] = av;
    }
  }
}

int enc_attn_bwd(const float* qkv, const float* dO, float* dqkv, int B, int n, int C, cudaStream_t stream) {
  RALD_REQUIRE(n > 0 && n <= 64 && C % 4 == 0 && C <= 256, "enc_attn_bwd: n=%d C=%d unsupported", n, C);
  enc_attn_bwd_kernel<<<dim3(B, 4), 256, 0, stream>>>(qkv, dO, dqkv, n, C, 1.0f / sqrtf((float)C));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

extern "C" {

int rald_gn_bwd(const float* x, const float* dy, const double* stats, const float* gamma, const float* beta, int B, int64_t V,
                int C, int groups, float eps, int swish, double* sums, const float* add, float* dx, void* stream) {
  return rald::gn_bwd(x, dy, stats, gamma, beta, B, V, C, groups, eps, swish, sums, add, dx,
                      static_cast<cudaStream_t>(stream));
}

int rald_enc_pad_transpose(const void* in, int in_f32, int B, int D, int H, int W, int C, int dil, int Wp, int copies,
                           int copy_rows, int pos_bias, void* out_t_bf16, int64_t ld, int panel_len, int halo, int n_panels,
                           double* colsum, void* stream) {
  return rald::enc_pad_transpose(in, in_f32, B, D, H, W, C, dil, Wp, copies, copy_rows, pos_bias, out_t_bf16, ld, panel_len,
                                 halo, n_panels, colsum, static_cast<cudaStream_t>(stream));
}

int rald_enc_stuff(const float* in, int B, int D, int H, int W, int C, void* out_bf16, void* stream) {
  return rald::enc_stuff(in, B, D, H, W, C, out_bf16, static_cast<cudaStream_t>(stream));
}

int rald_enc_attn_bwd(const float* qkv, const float* dO, float* dqkv, int B, int n, int C, void* stream) {
  return rald::enc_attn_bwd(qkv, dO, dqkv, B, n, C, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
