// SIMT kernels of the radar-cube encoder around the tcgen05 convolutions (all channels-last, [B, V, C]):
//   conv_in_kernel     3x3x3 Conv3d with 1..4 input channels (conv_in)        model/models_radar_encoder.py:161-163, 217
//   gn_stats_kernel    per (frame, group) sum / sum of squares                 GroupNorm(32, eps 1e-6) :9-12
//   gn_apply_kernel    y = swish?(GN(x)) -> bf16 operand of the next conv; or plain fp32 -> bf16 cast
//   enc_attn_kernel    single-head attention over <= 64 voxels, head_dim = C   AttnBlock :112-135
// These are HBM-bound passes (stats: 4 B/elem read; apply: 4 B read + 2 B written per element).
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

// ---------------------------------------------------------------------------------------------------
// conv_in: x [B, D, H, W, Cin] fp32 -> out [B, D, H, W, Cout] fp32, padding 1. One thread = one voxel x 64 output
// channels held in registers: per tap one input load (L1-resident neighbourhood) feeds 64 FMAs whose weights come
// from shared memory as warp-wide broadcasts, so the kernel runs at the FMA rate (1728 FMA per voxel for Cin = 1)
// instead of the load rate. Each thread writes its 256 contiguous output bytes.
// ---------------------------------------------------------------------------------------------------
constexpr int CIN_CO = 64;  // output channels per thread
__global__ void __launch_bounds__(128)
conv_in_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ out, int B, int D, int H, int W, int Cin, int Cout) {
  extern __shared__ float s_w[];  // [27*Cin][64] weights of this block's 64-channel slab (tap-major)
  const int co0 = blockIdx.y * CIN_CO;
  const int K = 27 * Cin;
  for (int i = threadIdx.x; i < K * CIN_CO; i += blockDim.x) {
    const int co = i % CIN_CO, k = i / CIN_CO;   // k = tap * Cin + ci
    const int tap = k / Cin, ci = k - tap * Cin;
    s_w[i] = w[((int64_t)(co0 + co) * Cin + ci) * 27 + tap];  // PyTorch layout [Cout][Cin][3][3][3]
  }
  __syncthreads();
  const int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)B * D * H * W;
  if (vox >= total) return;
  int64_t v = vox;
  const int wz = (int)(v % W); v /= W;
  const int hy = (int)(v % H); v /= H;
  const int dz = (int)(v % D);
  const int b = (int)(v / D);
  float acc[CIN_CO];
#pragma unroll
  for (int c = 0; c < CIN_CO; c += 4) {
    const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + co0 + c));
    acc[c] = bv.x; acc[c + 1] = bv.y; acc[c + 2] = bv.z; acc[c + 3] = bv.w;
  }
  for (int kd = 0; kd < 3; ++kd) {
    const int d = dz + kd - 1;
    if (d < 0 || d >= D) continue;
    for (int kh = 0; kh < 3; ++kh) {
      const int h = hy + kh - 1;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ww = wz + kw - 1;
        const bool ok = ww >= 0 && ww < W;   // predicated (not skipped): keeps the warp's broadcast loads uniform
        const float* xp = x + ((((int64_t)b * D + d) * H + h) * W + (ok ? ww : wz)) * Cin;
        const int tap = kd * 9 + kh * 3 + kw;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = ok ? __ldg(xp + ci) : 0.f;
          const float4* wr = reinterpret_cast<const float4*>(s_w + (tap * Cin + ci) * CIN_CO);
#pragma unroll
          for (int c = 0; c < CIN_CO / 4; ++c) {
            const float4 wv = wr[c];
            acc[4 * c + 0] = fmaf(xv, wv.x, acc[4 * c + 0]);
            acc[4 * c + 1] = fmaf(xv, wv.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(xv, wv.z, acc[4 * c + 2]);
            acc[4 * c + 3] = fmaf(xv, wv.w, acc[4 * c + 3]);
          }
        }
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(out + vox * Cout + co0);
#pragma unroll
  for (int c = 0; c < CIN_CO / 4; ++c) dst[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
}

// ---------------------------------------------------------------------------------------------------
// conv_in on the tensor cores (W a multiple of 16, one input channel — the shipped encoder): the FMA form above pays 1 728 FMAs per voxel and
// ran at 1.8 ms per 32 cubes against 0.33 ms for the mandatory write of the [B, V, 64] fp32 output. Here a warp owns 16
// consecutive voxels of a W row as the M side of mma.sync m16n8k16 products, K = (tap, input channel) padded to a
// multiple of 16, N = the 64 output channels of the slab. Both operands are split into bf16 hi + lo halves (hi*hi +
// hi*lo + lo*hi, fp32 accumulation: 16+ mantissa bits; the fp32 cube values are NOT rounded to bf16). A fragments are
// gathered straight from global memory / L1 (the taps a lane needs are the same for every tile: their offsets are
// decoded once), B fragments sit in shared memory split, packed and in fragment order (one conflict-free LDS.128 per
// three MMAs), output channel tiles are paired so that a lane owns four consecutive channels (16-byte stores).
// ---------------------------------------------------------------------------------------------------
template <int KSTEPS>
__global__ void __launch_bounds__(256, 2)
conv_in_mma_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out, int B, int D, int H, int W, int Cin, int Cout) {
  __shared__ __align__(16) uint32_t s_b[KSTEPS * 8 * 32 * 4];  // [k step][n tile][lane][b0 hi, b1 hi, b0 lo, b1 lo]
  __shared__ __align__(16) float s_bias[CIN_CO];
  const int co0 = blockIdx.y * CIN_CO;
  const int K = 27 * Cin;
  for (int i = threadIdx.x; i < KSTEPS * 16 * CIN_CO; i += blockDim.x) {
    const int c = i % CIN_CO, k = i / CIN_CO;          // k = tap * Cin + ci
    float wv = 0.f;
    if (k < K) {
      const int tap = k / Cin, ci = k - tap * Cin;
      wv = w[((int64_t)(co0 + c) * Cin + ci) * 27 + tap];  // PyTorch layout [Cout][Cin][3][3][3]
    }
    // channel c -> (n tile, column g): the 16 channels of a pair are dealt so that lane t of an accumulator owns channels
    // 4t .. 4t+3 of the pair: c % 16 = 4 (g / 2) + 2 ab + g % 2
    const int r = c & 15, nt = (c >> 4) * 2 + ((r >> 1) & 1), gg = (r >> 2) * 2 + (r & 1);
    const int ks = k >> 4, kk = k & 15, tt = (kk & 7) >> 1, which = kk >> 3;
    const __nv_bfloat16 hi = __float2bfloat16_rn(wv);
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(s_b + (((ks * 8 + nt) * 32 + gg * 4 + tt) << 2));
    dst[2 * which + (kk & 1)] = hi;
    dst[2 * (2 + which) + (kk & 1)] = __float2bfloat16_rn(wv - __bfloat162float(hi));
  }
  if (threadIdx.x < CIN_CO) s_bias[threadIdx.x] = bias[co0 + threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // this lane's K slots: k = 16 ks + {2t, 2t+1, 2t+8, 2t+9} -> offset of the tap's voxel from the output voxel
  int s_off[KSTEPS][4], s_tap[KSTEPS][4];   // s_tap: dd + 1 | (dh + 1) << 8 | (dw + 1) << 16; an invalid slot gets dd = 1 << 6
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = ks * 16 + 2 * t + (j & 1) + (j >> 1) * 8;
      const int tap = k / Cin, ci = k - tap * Cin;
      const int dd = tap / 9 - 1, dh = (tap / 3) % 3 - 1, dw = tap % 3 - 1;
      s_tap[ks][j] = (k < K ? dd + 1 : 64) | ((dh + 1) << 8) | ((dw + 1) << 16);
      s_off[ks][j] = ((dd * H + dh) * W + dw) * Cin + ci;
    }
  const int wtiles = W >> 4;
  const int64_t n_tiles = (int64_t)B * D * H * wtiles;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t tile = warp_global; tile < n_tiles; tile += warps_total) {
    int64_t v = tile;
    const int w0 = (int)(v % wtiles) << 4; v /= wtiles;
    const int hy = (int)(v % H); v /= H;
    const int dz = (int)(v % D);
    const int64_t vox0 = tile * 16;                       // first voxel of the tile ((b, d, h, w0) flattened)
    const float* xa = x + (vox0 + g) * Cin;               // row g; row g + 8 is 8 Cin further
    const int wa = w0 + g, wb = wa + 8;
    float acc[8][4];
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      const float4 b4 = reinterpret_cast<const float4*>(s_bias)[pr * 4 + t];   // channels 16 pr + 4t .. 4t+3
      acc[2 * pr][0] = acc[2 * pr][2] = b4.x;
      acc[2 * pr][1] = acc[2 * pr][3] = b4.y;
      acc[2 * pr + 1][0] = acc[2 * pr + 1][2] = b4.z;
      acc[2 * pr + 1][1] = acc[2 * pr + 1][3] = b4.w;
    }
    float va[KSTEPS][4], vb[KSTEPS][4];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = s_tap[ks][j];
        const bool in = (unsigned)(dz + (m & 255) - 1) < (unsigned)D && (unsigned)(hy + ((m >> 8) & 255) - 1) < (unsigned)H;
        const int dw = (m >> 16) - 1;
        va[ks][j] = (in && (unsigned)(wa + dw) < (unsigned)W) ? __ldg(xa + s_off[ks][j]) : 0.f;
        vb[ks][j] = (in && (unsigned)(wb + dw) < (unsigned)W) ? __ldg(xa + 8 * Cin + s_off[ks][j]) : 0.f;
      }
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      uint32_t ahi[4], alo[4];
      split_pair(va[ks][0], va[ks][1], ahi[0], alo[0]);   // a0: row g,     k = 2t, 2t+1
      split_pair(vb[ks][0], vb[ks][1], ahi[1], alo[1]);   // a1: row g + 8
      split_pair(va[ks][2], va[ks][3], ahi[2], alo[2]);   // a2: row g,     k = 2t+8, 2t+9
      split_pair(vb[ks][2], vb[ks][3], ahi[3], alo[3]);   // a3: row g + 8
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint4 bw[4];
#pragma unroll
        for (int n = 0; n < 4; ++n) bw[n] = reinterpret_cast<const uint4*>(s_b)[(ks * 8 + half * 4 + n) * 32 + lane];
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_bf16(acc[half * 4 + n], alo, bw[n].x, bw[n].y);
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_bf16(acc[half * 4 + n], ahi, bw[n].z, bw[n].w);
#pragma unroll
        for (int n = 0; n < 4; ++n) mma_bf16(acc[half * 4 + n], ahi, bw[n].x, bw[n].y);
      }
    }
    float4* oa = reinterpret_cast<float4*>(out + (vox0 + g) * Cout + co0) + t;
    float4* ob = reinterpret_cast<float4*>(out + (vox0 + g + 8) * Cout + co0) + t;
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      oa[pr * 4] = make_float4(acc[2 * pr][0], acc[2 * pr][1], acc[2 * pr + 1][0], acc[2 * pr + 1][1]);
      ob[pr * 4] = make_float4(acc[2 * pr][2], acc[2 * pr][3], acc[2 * pr + 1][2], acc[2 * pr + 1][3]);
    }
  }
}

template <int KSTEPS>
static void launch_conv_in_mma(const float* x, const float* w, const float* bias, float* out, int B, int D, int H, int W,
                               int Cin, int Cout, cudaStream_t stream) {
  const int64_t tiles = (int64_t)B * D * H * (W / 16);
  int64_t blocks = (tiles + 7) / 8;
  const int64_t cap = (int64_t)device_sm_count() * 4;   // persistent warps: the weight slab is staged once per CTA
  if (blocks > cap) blocks = cap;
  conv_in_mma_kernel<KSTEPS><<<dim3((unsigned)blocks, (unsigned)(Cout / CIN_CO)), 256, 0, stream>>>(x, w, bias, out, B, D, H,
                                                                                                  W, Cin, Cout);
}

int enc_conv_in(const float* x, const float* w, const float* bias, float* out, int B, int D, int H, int W, int Cin,
                int Cout, cudaStream_t stream) {
  RALD_REQUIRE(Cin >= 1 && Cin <= 4, "conv_in: Cin=%d must be in [1, 4]", Cin);
  RALD_REQUIRE(Cout % CIN_CO == 0, "conv_in: Cout=%d must be a multiple of %d", Cout, CIN_CO);
  const int64_t total = (int64_t)B * D * H * W;
  ProfScope prof(FAM_OTHER, stream, (double)total * Cout * 4.0);
  if (W % 16 == 0 && 27 * Cin <= 32) {
    // the shipped encoder (one input channel): K = 27 taps padded to two k steps of 16
    launch_conv_in_mma<2>(x, w, bias, out, B, D, H, W, Cin, Cout, stream);
  } else {
    // other widths / more input channels: one thread per voxel on the FMA pipe
    const int smem = 27 * Cin * CIN_CO * sizeof(float);
    dim3 grid((unsigned)((total + 127) / 128), (unsigned)(Cout / CIN_CO));
    conv_in_kernel<<<grid, 128, smem, stream>>>(x, w, bias, out, B, D, H, W, Cin, Cout);
  }
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm statistics: stats[b][g] = {sum, sum of squares} (double, accumulated with atomics; zeroed here by a
// memset node issued before the kernel). grid = (chunks, B).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_stats_kernel(const float* __restrict__ x, int64_t V, int C, int groups, int64_t vox_per_block,
                double* __restrict__ stats) {
  // per-thread partials [voxel lane][channel], reduced in a FIXED order (no floating-point atomics inside the block:
  // the block's contribution is bit-reproducible; only the fp64 atomics across blocks remain order-dependent, at
  // 1e-16 relative)
  __shared__ float s_sum[256 * 4];
  __shared__ float s_sq[256 * 4];
  const int quads = C >> 2;             // float4 columns per voxel (16 / 32 / 64)
  const int vlanes = blockDim.x / quads;
  const int cq = threadIdx.x % quads;
  const int vl = threadIdx.x / quads;
  const int b = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * vox_per_block;
  const int64_t v1 = (v0 + vox_per_block) < V ? (v0 + vox_per_block) : V;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (vl < vlanes) {
    const float4* xb = reinterpret_cast<const float4*>(x + (int64_t)b * V * C);
    for (int64_t v = v0 + vl; v < v1; v += vlanes) {
      const float4 a = xb[v * quads + cq];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x = fmaf(a.x, a.x, q.x); q.y = fmaf(a.y, a.y, q.y); q.z = fmaf(a.z, a.z, q.z); q.w = fmaf(a.w, a.w, q.w);
    }
    reinterpret_cast<float4*>(s_sum)[vl * quads + cq] = s;   // s_sum[vl][4*cq .. 4*cq+3]
    reinterpret_cast<float4*>(s_sq)[vl * quads + cq] = q;
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int cpg = C / groups;
    double gs = 0.0, gq = 0.0;
    for (int l = 0; l < vlanes; ++l) {
      for (int j = 0; j < cpg; ++j) {
        gs += (double)s_sum[l * C + threadIdx.x * cpg + j];
        gq += (double)s_sq[l * C + threadIdx.x * cpg + j];
      }
    }
    atomicAdd(&stats[((int64_t)b * groups + threadIdx.x) * 2 + 0], gs);
    atomicAdd(&stats[((int64_t)b * groups + threadIdx.x) * 2 + 1], gq);
  }
}

int gn_stats(const float* x, int B, int64_t V, int C, int groups, double* stats, cudaStream_t stream) {
  RALD_REQUIRE(C % 4 == 0 && C <= 256 && C >= 4 && 256 % (C / 4) == 0, "gn_stats: C=%d unsupported", C);
  RALD_REQUIRE(groups > 0 && groups <= 256 && C % groups == 0, "gn_stats: groups=%d does not divide C=%d", groups, C);
  RALD_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * B * groups, stream));
  // ~16 K elements per thread block keeps >= 2 waves at the large levels and one block at the small ones
  int64_t vox_per_block = (16384 * 4) / C;
  if (vox_per_block < 1) vox_per_block = 1;
  const int64_t chunks = (V + vox_per_block - 1) / vox_per_block;
  dim3 grid((unsigned)chunks, (unsigned)B);
  ProfScope prof(FAM_GN, stream, (double)B * V * C * 4.0);
  gn_stats_kernel<<<grid, 256, 0, stream>>>(x, V, C, groups, vox_per_block, stats);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// GroupNorm apply: mode 0 = GN + swish, 1 = GN only, 2 = cast only (stats/gamma/beta unused). Output bf16.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, int64_t V, int C, int groups,
                float eps, int mode, int64_t quads_per_block) {
  __shared__ float s_a[256];
  __shared__ float s_d[256];
  const int b = blockIdx.y;
  if (mode != 2) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int cpg = C / groups;
      const int g = c / cpg;
      const double n = (double)V * cpg;
      const double mean = stats[((int64_t)b * groups + g) * 2 + 0] / n;
      double var = stats[((int64_t)b * groups + g) * 2 + 1] / n - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      const float a = rstd * gamma[c];
      s_a[c] = a;
      s_d[c] = beta[c] - (float)mean * a;
    }
    __syncthreads();
  }
  const int quads = C >> 2;
  const int64_t nq = V * quads;  // float4s of this frame
  const int64_t q0 = (int64_t)blockIdx.x * quads_per_block;
  const int64_t q1 = (q0 + quads_per_block) < nq ? (q0 + quads_per_block) : nq;
  const float4* xb = reinterpret_cast<const float4*>(x + (int64_t)b * V * C);
  uint2* ob = reinterpret_cast<uint2*>(out + (int64_t)b * V * C);
  for (int64_t i = q0 + threadIdx.x; i < q1; i += blockDim.x) {
    float4 v = xb[i];
    if (mode != 2) {
      const int c = (int)(i % quads) * 4;
      v.x = fmaf(v.x, s_a[c + 0], s_d[c + 0]);
      v.y = fmaf(v.y, s_a[c + 1], s_d[c + 1]);
      v.z = fmaf(v.z, s_a[c + 2], s_d[c + 2]);
      v.w = fmaf(v.w, s_a[c + 3], s_d[c + 3]);
      if (mode == 0) {
        v.x = v.x / (1.0f + __expf(-v.x));
        v.y = v.y / (1.0f + __expf(-v.y));
        v.z = v.z / (1.0f + __expf(-v.z));
        v.w = v.w / (1.0f + __expf(-v.w));
      }
    }
    ob[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

int gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* out_bf16, int B,
             int64_t V, int C, int groups, float eps, int mode, cudaStream_t stream) {
  RALD_REQUIRE(C % 4 == 0 && C <= 256, "gn_apply: C=%d unsupported", C);
  RALD_REQUIRE(mode >= 0 && mode <= 2, "gn_apply: mode %d", mode);
  RALD_REQUIRE(mode == 2 || (groups > 0 && C % groups == 0), "gn_apply: groups=%d does not divide C=%d", groups, C);
  const int64_t nq = V * (C / 4);
  const int64_t quads_per_block = 256 * 16;
  const int64_t chunks = (nq + quads_per_block - 1) / quads_per_block;
  dim3 grid((unsigned)chunks, (unsigned)B);
  ProfScope prof(FAM_GN, stream, (double)B * V * C * 6.0);
  gn_apply_kernel<<<grid, 256, 0, stream>>>(x, stats, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out_bf16), V, C,
                                            groups == 0 ? 1 : groups, eps, mode, quads_per_block);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// AttnBlock core: qkv fp32 [B*n, 3C] (q | k | v), one head of width C over the n <= 64 voxels of a frame.
// out bf16 [B*n, C] = softmax(q k^T / sqrt(C)) v. One CTA per frame; everything in shared memory, fp32.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
enc_attn_kernel(const float* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int n, int C, float scale) {
  extern __shared__ float sm[];
  const int ldc = C + 4;  // pad: rows start on different banks
  float* s_q = sm;
  float* s_k = s_q + 64 * ldc;
  float* s_v = s_k + 64 * ldc;
  float* s_p = s_v + 64 * ldc;  // [64][65]
  const int b = blockIdx.x;
  const float* base = qkv + (int64_t)b * n * 3 * C;
  const int quads = C >> 2;
  for (int i = threadIdx.x; i < n * 3 * quads; i += blockDim.x) {
    const int row = i / (3 * quads);
    const int col4 = i - row * 3 * quads;
    const float4 v = reinterpret_cast<const float4*>(base + (int64_t)row * 3 * C)[col4];
    const int which = col4 / quads, c = (col4 - which * quads) * 4;
    float* dst = (which == 0 ? s_q : which == 1 ? s_k : s_v) + row * ldc + c;
    *reinterpret_cast<float4*>(dst) = v;
  }
  __syncthreads();
  // scores: thread (i, jq) computes 16 columns of row i
  {
    const int i = threadIdx.x >> 2, jq = threadIdx.x & 3;
    if (i < n) {
      for (int jj = 0; jj < 16; ++jj) {
        const int j = jq * 16 + jj;
        float acc = 0.f;
        if (j < n) {
          const float4* qa = reinterpret_cast<const float4*>(s_q + i * ldc);
          const float4* kb = reinterpret_cast<const float4*>(s_k + j * ldc);
          for (int c = 0; c < quads; ++c) {
            const float4 a = qa[c], k4 = kb[c];
            acc = fmaf(a.x, k4.x, acc); acc = fmaf(a.y, k4.y, acc);
            acc = fmaf(a.z, k4.z, acc); acc = fmaf(a.w, k4.w, acc);
          }
          acc *= scale;
        } else {
          acc = -INFINITY;
        }
        s_p[i * 65 + j] = acc;
      }
    }
  }
  __syncthreads();
  // softmax over j: 4 threads per row
  {
    const int i = threadIdx.x >> 2, jq = threadIdx.x & 3;
    const bool ok = i < n;  // rows >= n hold no data; every lane still takes part in the shuffles
    float e[16];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      e[jj] = ok ? s_p[i * 65 + jq * 16 + jj] : 0.f;
      mx = fmaxf(mx, e[jj]);
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      e[jj] = expf(e[jj] - mx);
      sum += e[jj];
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const float inv = 1.0f / sum;
    if (ok) {
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) s_p[i * 65 + jq * 16 + jj] = e[jj] * inv;
    }
  }
  __syncthreads();
  // out[i][c] = sum_j p[i][j] v[j][c]; thread handles channel c for 64/(256/C)... generic strided loop
  for (int idx = threadIdx.x; idx < n * C; idx += blockDim.x) {
    const int i = idx / C, c = idx - i * C;
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = fmaf(s_p[i * 65 + j], s_v[j * ldc + c], acc);
    out[((int64_t)b * n + i) * C + c] = __float2bfloat16_rn(acc);
  }
}

int enc_attn(const float* qkv, void* out_bf16, int B, int n, int C, cudaStream_t stream) {
  RALD_REQUIRE(n >= 1 && n <= 64, "enc_attn: %d voxels per frame (the kernel handles <= 64)", n);
  RALD_REQUIRE(C % 4 == 0 && C <= 256, "enc_attn: C=%d unsupported", C);
  const int smem = (3 * 64 * (C + 4) + 64 * 65) * sizeof(float);
  static int configured = 48 * 1024;
  if (smem > configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(enc_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  enc_attn_kernel<<<B, 256, smem, stream>>>(qkv, reinterpret_cast<__nv_bfloat16*>(out_bf16), n, C,
                                            1.0f / sqrtf((float)C));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald
