// Small fp32 kernels around the denoiser blocks (latency / HBM bound):
//   temb_kernel      PositionalEmbedding + map_layer0/1 + SiLU          models_radar_generation.py:27-33, 217-219
//   adaln_kernel     all depth*3 AdaLayerNorm linears for S sigmas      models_radar_generation.py:128-129
//   boundary_kernel  final LayerNorm + proj_out + EDM preconditioning + Euler/Heun update + next proj_in
//                    (16-row tiles, split-bf16 mma.sync products with fp32 accumulation, fp32 update arithmetic)
//                                                                        :230-232, :422-429, :265-273, :221
//   radar_tokens_kernel  token projection + r/a/e embeddings            :390-405
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------
// t_emb[s, :] = silu(W1 silu(W0 [cos(c f), sin(c f)] + b0) + b1),  c = ln(sigma_s) / 4
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
temb_kernel(const float* __restrict__ sigma, const float* __restrict__ freqs, int half, const float* __restrict__ w0,
            const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ b1, int dim,
            float* __restrict__ t_emb) {
  extern __shared__ float sm[];
  float* emb = sm;             // [2*half]
  float* h0 = sm + 2 * half;   // [dim]
  const int s = blockIdx.x;
  const float c_noise = logf(sigma[s]) / 4.0f;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float x = c_noise * freqs[j];
    emb[j] = cosf(x);
    emb[half + j] = sinf(x);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int in0 = 2 * half;
  for (int r = warp; r < dim; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < in0; k += 32) acc += w0[(int64_t)r * in0 + k] * emb[k];
    acc = warp_sum(acc);
    if (lane == 0) h0[r] = silu(acc + b0[r]);
  }
  __syncthreads();
  for (int r = warp; r < dim; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc += w1[(int64_t)r * dim + k] * h0[k];
    acc = warp_sum(acc);
    if (lane == 0) t_emb[(int64_t)s * dim + r] = silu(acc + b1[r]);
  }
}

// mod[s, r] = ada_w[r, :] . t_emb[s, :] + ada_b[r]   for r < R = depth*3*2*dim ; one warp per row r.
constexpr int ADA_SCHUNK = 16;
__global__ void __launch_bounds__(256)
adaln_kernel(const float* __restrict__ t_emb, int S, const float* __restrict__ ada_w, const float* __restrict__ ada_b,
             int64_t R, float* __restrict__ mod) {
  __shared__ float4 te[ADA_SCHUNK][128];
  const int s0 = blockIdx.y * ADA_SCHUNK;
  const int ns = (S - s0) < ADA_SCHUNK ? (S - s0) : ADA_SCHUNK;
  for (int i = threadIdx.x; i < ns * 128; i += blockDim.x)
    te[i / 128][i % 128] = reinterpret_cast<const float4*>(t_emb + (int64_t)(s0 + i / 128) * 512)[i % 128];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= R) return;
  float4 w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = __ldg(reinterpret_cast<const float4*>(ada_w + r * 512) + j * 32 + lane);
  const float bias = ada_b[r];
  for (int s = 0; s < ns; ++s) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = te[s][j * 32 + lane];
      acc += (w[j].x * t.x + w[j].y * t.y) + (w[j].z * t.z + w[j].w * t.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) mod[(int64_t)(s0 + s) * R + r] = acc + bias;
  }
}

int dit_mod_table(const float* sigma, int S, const float* freqs, int half, const float* map0_w, const float* map0_b,
                  const float* map1_w, const float* map1_b, const float* ada_w, const float* ada_b, int depth,
                  int dim, float* t_emb_ws, float* mod, cudaStream_t stream) {
  RALD_REQUIRE(dim == 512, "dit_mod_table: dim=%d unsupported (512 only)", dim);
  RALD_REQUIRE(S > 0 && depth > 0, "dit_mod_table: bad sizes");
  const int smem = (2 * half + dim) * sizeof(float);
  temb_kernel<<<S, 512, smem, stream>>>(sigma, freqs, half, map0_w, map0_b, map1_w, map1_b, dim, t_emb_ws);
  RALD_LAUNCHED();
  const int64_t R = (int64_t)depth * 3 * 2 * dim;
  dim3 grid((unsigned)((R + 7) / 8), (unsigned)((S + ADA_SCHUNK - 1) / ADA_SCHUNK));
  adaln_kernel<<<grid, 256, 0, stream>>>(t_emb_ws, S, ada_w, ada_b, R, mod);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// boundary kernel: tiles of 16 latent rows (dim = 512, channels <= 32) on the tensor cores
// ---------------------------------------------------------------------------------------------------
struct BoundaryParams {
  const float* h;          // [T, 512] residual stream after the last block (null when mode == 3)
  const float* ln_w;       // [512]
  const float* ln_b;       // [512]
  const float* w_out_t;    // [512][32]  proj_out transposed, zero-padded to 32 channels
  const float* w_in_t;     // [C][512]   proj_in transposed
  const float* x_in;       // [T, C] un-scaled network input of this evaluation
  const float* x_base;     // [T, C] x_hat of the current step (mode 2)
  float* d_buf;            // [T, C] d_cur (written in mode 1, read in mode 2)
  float* x_out;            // [T, C]
  float* h_next;           // [T, 512] or null
  const float* sigma;      // sigma of this evaluation (per frame with stride, or shared with stride 0)
  const float* sigma_other;
  int64_t sigma_stride, sigma_other_stride;
  int mode;                // 0 = D only, 1 = Euler, 2 = Heun, 3 = init (x_out = x_in * sigma), 4 = project only
  int rows_per_frame, C;
  int64_t T;
  float sigma_data;
  const uint32_t* pack;    // BD_PACK_WORDS words: the weights split, packed and in fragment order (boundary_pack_kernel)
  int pack_fresh;          // the pack was written by the preceding kernel of the stream: fetch it after the grid dependency
};

// Both projections are 512 <-> 32 GEMMs with two or three flops per byte of h: as shared-memory GEMVs on the FMA pipe
// they ran at the shared-memory bandwidth (0.13 of the HBM roofline at 32 768 rows, 35 us for 4 096 rows). Here a
// GROUP of four warps owns a tile of 16 rows and runs them as mma.sync m16n8k16 products of SPLIT-bf16 operands
// (x = hi + lo, two bf16 halves = 16+ mantissa bits; hi*hi + hi*lo + lo*hi in one fp32 accumulator: products exact to
// ~2^-16, two orders below the bf16 rounding of the network that produced h). tools/micro/mma_sync_rate.cu: the legacy
// tensor path issues one m16n8k16 (or one TF32 m16n8k8) per 2 clocks per SM — split TF32 (21 bits) needs twice the
// instructions and was MMA-bound at 60 us per 32 768 rows; the bf16 split keeps the tensor time under the HBM time.
//   phase 1  warp q streams columns [128q, 128q+128) of the 16 rows of h STRAIGHT FROM GLOBAL MEMORY INTO A FRAGMENTS
//            (the k index of an MMA is a free permutation: lane (g, t) takes 8 consecutive floats of rows g and g+8,
//            so a quad reads 128 contiguous bytes of a row) and accumulates (h - K) W' for the four 8-channel tiles,
//            W' = diag(ln_w) W_out^T, K = a per-row shift (median of three samples of the row) that keeps the
//            one-pass moments and the split operands well conditioned; the same pass takes sum(h - K) and sum((h - K)^2) in fp32.
//            LayerNorm is applied algebraically: F = rstd (acc - mean' colsum(W')) + ln_b W_out^T.
//            The loads of the NEXT tile are issued right after this phase, under the reduction and phases 2 / 3.
//   reduce   the four partial accumulators / moments meet in shared memory; warp q sums (fixed order) channel tile q.
//   phase 2  warp q: EDM preconditioning + Euler / Heun update on its accumulator fragment (rows g, g+8; channels
//            8 q + 2t, +1), the fp32 arithmetic of the reference (:422-429, :265-273); c_in x_next goes back through
//            shared memory already split and packed — in the layout in which it IS the A fragment of the next product.
//   phase 3  warp q: h_next[:, 128q : 128q+128] = (c_in x_next) W_in, n tiles paired so that a lane owns four
//            consecutive output columns (16-byte stores, a quad writes 64 contiguous bytes per row).
// Every row's arithmetic is independent of the tile it sits in and of the batch: a frame computes bit-identical values
// alone and in a batch. Weights sit in shared memory split, packed and in fragment order (hi and lo halves of an
// element take the 32 bits the fp32 value took: two conflict-free LDS.128 per 6 MMAs).
constexpr int BD_MAX_GROUPS = 4;             // 16-row tiles in flight per CTA (four warps each)
constexpr int BD_PART = 20;                  // 16 accumulators + 4 moments per lane
constexpr int BD_PACK_WORDS = 16384 + 16384 + 64;   // W' | W_in | column sums of W' | ln_b W_out^T
constexpr int BD_SMEM_WORDS = BD_PACK_WORDS + BD_MAX_GROUPS * 4 * BD_PART * 32 + BD_MAX_GROUPS * 16 * 32 + 4;
// word index of element (k, n) of W' = diag(ln_w) W_out^T: [kb = k/32][nt = n/8][hi | lo][lane = (n%8) 4 + (k%32)/8][e/2],
// half e%2, e = k%8 — a lane's four hi words (then its four lo words) are one conflict-free LDS.128
__device__ __forceinline__ int bd_wout_word(int k, int n, int hl) {
  return (((((k >> 5) * 4 + (n >> 3)) * 2 + hl) * 32 + (n & 7) * 4 + ((k & 31) >> 3)) << 2) + ((k & 7) >> 1);
}
// word index of element (c, col) of W_in^T: [pair = col/16][ab][hi | lo][lane = g 4 + (c%8)/2][ks = c/8], half c%2, where
// the 16 columns of a pair are dealt so that lane t of an accumulator owns columns 4t .. 4t+3: col%16 = 4 (g/2) + 2 ab + g%2
__device__ __forceinline__ int bd_win_word(int c, int col, int hl) {
  const int r = col & 15, ab = (r >> 1) & 1, gg = (r >> 2) * 2 + (r & 1);
  return 16384 + ((((((col >> 4) * 2 + ab) * 2 + hl) * 32 + gg * 4 + ((c & 7) >> 1)) << 2) + (c >> 3));
}

__device__ __forceinline__ void group_barrier(int grp) {
  asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
}
// Weights -> the image the boundary kernel copies into shared memory in one bulk transfer. One CTA; runs once per weight
// version (rald_dit_boundary_pack) or in front of a launch that brings raw weights (rald_dit_boundary).
__global__ void __launch_bounds__(512)
boundary_pack_kernel(const float* __restrict__ ln_w, const float* __restrict__ ln_b, const float* __restrict__ w_out_t,
                     const float* __restrict__ w_in_t, int C, uint32_t* __restrict__ pack) {
  __shared__ float s_red[2][64][32];
  __nv_bfloat16* halves = reinterpret_cast<__nv_bfloat16*>(pack);
  const int tid = threadIdx.x;
  // W' and its column sums (fixed order: thread partials over k = tid/8 + 64 i, then the 64 partials of a channel)
  const int k0 = tid >> 3, n4 = (tid & 7) * 4;
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, bw[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < 8; ++i) {
    const int k = k0 + 64 * i;
    const float4 w = __ldg(reinterpret_cast<const float4*>(w_out_t) + tid + 512 * i);
    const float lw = __ldg(ln_w + k), lb = __ldg(ln_b + k);
    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float wp = wv[c] * lw;
      const __nv_bfloat16 hi = __float2bfloat16_rn(wp);
      halves[2 * bd_wout_word(k, n4 + c, 0) + (k & 1)] = hi;
      halves[2 * bd_wout_word(k, n4 + c, 1) + (k & 1)] = __float2bfloat16_rn(wp - __bfloat162float(hi));
      cs[c] += wp;
      bw[c] = fmaf(lb, wv[c], bw[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    s_red[0][k0][n4 + c] = cs[c];
    s_red[1][k0][n4 + c] = bw[c];
  }
  for (int i = tid; i < 32 * 512; i += 512) {
    const int c = i >> 9, col = i & 511;
    const float w = c < C ? __ldg(w_in_t + (int64_t)c * 512 + col) : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
    halves[2 * bd_win_word(c, col, 0) + (c & 1)] = hi;
    halves[2 * bd_win_word(c, col, 1) + (c & 1)] = __float2bfloat16_rn(w - __bfloat162float(hi));
  }
  __syncthreads();
  if (tid < 64) {
    float acc = 0.f;
    for (int j = 0; j < 64; ++j) acc += s_red[tid >> 5][j][tid & 31];
    reinterpret_cast<float*>(pack)[32768 + tid] = acc;   // [32768, 32800): column sums, [32800, 32832): ln_b W_out^T
  }
}

struct BoundaryRows {   // one warp's share of a tile's h rows, in flight
  float4 a[4][2], b[4][2];
  float ka[3], kb[3];   // three samples of each row for the shift
};
// columns the per-row shift is taken from (any three; spread over the row)
constexpr int BD_K0 = 5, BD_K1 = 173, BD_K2 = 347;
__device__ __forceinline__ float median3(float a, float b, float c) {
  return fmaxf(fminf(a, b), fminf(fmaxf(a, b), c));
}
__device__ __forceinline__ void boundary_load_rows(BoundaryRows& v, const float* h, int64_t tile, int g, int t, int q) {
  const float4* pa = reinterpret_cast<const float4*>(h) + (tile * 16 + g) * 128;   // row g of the tile
  const float4* hA = pa + q * 32 + 2 * t;                                           // this lane's 32 bytes of block 4 q
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v.a[i][0] = hA[i * 8];
    v.a[i][1] = hA[i * 8 + 1];
    v.b[i][0] = hA[8 * 128 + i * 8];      // row g + 8
    v.b[i][1] = hA[8 * 128 + i * 8 + 1];
  }
  const float* fa = reinterpret_cast<const float*>(pa);
  v.ka[0] = fa[BD_K0]; v.ka[1] = fa[BD_K1]; v.ka[2] = fa[BD_K2];
  v.kb[0] = fa[8 * 512 + BD_K0]; v.kb[1] = fa[8 * 512 + BD_K1]; v.kb[2] = fa[8 * 512 + BD_K2];
}

template <int BD_GROUPS, int MODE, bool NEXT>
__global__ void __launch_bounds__(BD_GROUPS * 128, 1)
boundary_kernel(const BoundaryParams p) {
  extern __shared__ __align__(128) uint32_t smw[];
  uint32_t* s_wout = smw;              // the pack: W' ...
  uint32_t* s_win = smw;               // ... W_in (bd_win_word carries its offset) ...
  const float* s_cs = reinterpret_cast<const float*>(smw + 32768);  // ... [32] column sums of W', [32] ln_b W_out^T
  const float* s_bw = s_cs + 32;
  float* s_part = reinterpret_cast<float*>(smw + BD_PACK_WORDS);    // [group][warp][BD_PART][lane]
  uint32_t* s_xs = reinterpret_cast<uint32_t*>(s_part + BD_MAX_GROUPS * 4 * BD_PART * 32);  // [group][hi/lo 2][row 2][nt 4][lane]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_xs + BD_MAX_GROUPS * 16 * 32);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = warp >> 2, q = warp & 3, g = lane >> 2, t = lane & 3;
  constexpr bool need_net = MODE < 3;
  constexpr bool need_next = NEXT;

  // ---- the weight image: one elected thread, bulk copies straight into shared memory (no thread touches a weight).
  // A pack made at weight-packing time is a constant: it is fetched while the preceding kernel is still running ----
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_barrier_init();
  }
  if (p.pack_fresh) pdl_wait();
  if (tid == 0) {
    constexpr uint32_t kBytes = BD_PACK_WORDS * 4, kChunk = 32768;
    mbar_arrive_expect_tx(s_bar, kBytes);
    for (uint32_t off = 0; off < kBytes; off += kChunk) {
      const uint32_t n = kBytes - off < kChunk ? kBytes - off : kChunk;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(reinterpret_cast<const char*>(smw) + off)),
                   "l"(reinterpret_cast<const char*>(p.pack) + off), "r"(n), "r"(smem_u32(s_bar))
                   : "memory");
    }
  }
  if (!p.pack_fresh) pdl_wait();  // h / x / d come from the preceding kernels
  pdl_launch_dependents();
  __syncthreads();                // the barrier's initialisation is visible to every waiter
  mbar_wait(s_bar, 0);
  float my_cs[2], my_bw[2];       // of this warp's channels 8 q + 2 t + b
#pragma unroll
  for (int b = 0; b < 2; ++b) {
    my_cs[b] = s_cs[q * 8 + 2 * t + b];
    my_bw[b] = s_bw[q * 8 + 2 * t + b];
  }

  const float sd = p.sigma_data;
  const int64_t n_tiles = p.T >> 4;
  const int64_t tile_step = (int64_t)gridDim.x * BD_GROUPS;
  float* my_part = s_part + ((grp * 4 + q) * BD_PART) * 32 + lane;
  uint32_t* my_xs = s_xs + grp * 16 * 32 + lane;

  int64_t tile = blockIdx.x + (int64_t)gridDim.x * grp;
  BoundaryRows v;
  if (need_net && tile < n_tiles) boundary_load_rows(v, p.h, tile, g, t, q);

  for (; tile < n_tiles; tile += tile_step) {
    const int64_t rowA = tile * 16 + g, rowB = rowA + 8;
    // ---- this warp's [T, C] operands of phase 2 (rows A / B, channels 8 q + 2 t + b), requested ahead of phase 1 ----
    float xin[2][2], dcur[2][2], xbase[2][2];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ch = q * 8 + 2 * t + b;
        const bool ok = ch < p.C;
        const int64_t o = (r == 0 ? rowA : rowB) * p.C + ch;
        xin[r][b] = ok ? p.x_in[o] : 0.f;
        dcur[r][b] = (ok && MODE == 2) ? p.d_buf[o] : 0.f;
        xbase[r][b] = (ok && MODE == 2) ? p.x_base[o] : 0.f;
      }
    if (need_net) {
      // ---- phase 1: this warp's 128 columns of the 16 rows ----
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      float s1A = 0.f, s2A = 0.f, s1B = 0.f, s2B = 0.f;
      // the shift only has to sit inside the bulk of the row (the split operands carry 16+ bits RELATIVE TO |h - K|): the
      // median of three samples ignores a massive-activation channel among them (tests/test_cpu_boundary_numerics.py)
      const float kA = median3(v.ka[0], v.ka[1], v.ka[2]);
      const float kB = median3(v.kb[0], v.kb[1], v.kb[2]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kb = q * 4 + i;
        const float ra[8] = {v.a[i][0].x - kA, v.a[i][0].y - kA, v.a[i][0].z - kA, v.a[i][0].w - kA,
                             v.a[i][1].x - kA, v.a[i][1].y - kA, v.a[i][1].z - kA, v.a[i][1].w - kA};
        const float rb[8] = {v.b[i][0].x - kB, v.b[i][0].y - kB, v.b[i][0].z - kB, v.b[i][0].w - kB,
                             v.b[i][1].x - kB, v.b[i][1].y - kB, v.b[i][1].z - kB, v.b[i][1].w - kB};
        uint32_t ahi[2][4], alo[2][4];   // [k16 step j][a0 .. a3]: elements 4j, 4j+1 | 4j+2, 4j+3 of rows A / B
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          split_pair(ra[4 * j], ra[4 * j + 1], ahi[j][0], alo[j][0]);
          split_pair(rb[4 * j], rb[4 * j + 1], ahi[j][1], alo[j][1]);
          split_pair(ra[4 * j + 2], ra[4 * j + 3], ahi[j][2], alo[j][2]);
          split_pair(rb[4 * j + 2], rb[4 * j + 3], ahi[j][3], alo[j][3]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          s1A += ra[e];
          s2A = fmaf(ra[e], ra[e], s2A);
          s1B += rb[e];
          s2B = fmaf(rb[e], rb[e], s2B);
        }
        // consecutive MMAs go to different accumulators (an accumulator's next MMA is four instructions away)
        uint4 whi[4], wlo[4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint4* wp = reinterpret_cast<const uint4*>(s_wout) + (kb * 4 + nt) * 64 + lane;
          whi[nt] = wp[0];
          wlo[nt] = wp[32];
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], alo[0], whi[nt].x, whi[nt].y);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], ahi[0], wlo[nt].x, wlo[nt].y);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], ahi[0], whi[nt].x, whi[nt].y);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], alo[1], whi[nt].z, whi[nt].w);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], ahi[1], wlo[nt].z, wlo[nt].w);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_bf16(acc[nt], ahi[1], whi[nt].z, whi[nt].w);
      }
      // ---- the next tile's rows start their way in under the reduction, phase 2 and phase 3 ----
      if (tile + tile_step < n_tiles) boundary_load_rows(v, p.h, tile + tile_step, g, t, q);
      s1A += __shfl_xor_sync(0xffffffffu, s1A, 1);
      s2A += __shfl_xor_sync(0xffffffffu, s2A, 1);
      s1B += __shfl_xor_sync(0xffffffffu, s1B, 1);
      s2B += __shfl_xor_sync(0xffffffffu, s2B, 1);
      s1A += __shfl_xor_sync(0xffffffffu, s1A, 2);
      s2A += __shfl_xor_sync(0xffffffffu, s2A, 2);
      s1B += __shfl_xor_sync(0xffffffffu, s1B, 2);
      s2B += __shfl_xor_sync(0xffffffffu, s2B, 2);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int c = 0; c < 4; ++c) my_part[(nt * 4 + c) * 32] = acc[nt][c];
      my_part[16 * 32] = s1A;
      my_part[17 * 32] = s2A;
      my_part[18 * 32] = s1B;
      my_part[19 * 32] = s2B;
    }
    group_barrier(grp);
    {
      // ---- reduce (fixed order: warp 0, 1, 2, 3) + phase 2: warp q owns channel tile q (channels 8 q + 2 t + b) ----
      float F[2][2] = {{0.f, 0.f}, {0.f, 0.f}};  // [row A / B][b]
      if (need_net) {
        const float* pp = s_part + (grp * 4 * BD_PART) * 32 + lane;
        float tot[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int j = c < 4 ? q * 4 + c : 12 + c;   // this tile's four accumulators, then the four moments
          tot[c] = ((pp[j * 32] + pp[(BD_PART + j) * 32]) + pp[(2 * BD_PART + j) * 32]) + pp[(3 * BD_PART + j) * 32];
        }
        const float mA = tot[4] * (1.0f / 512), mB = tot[6] * (1.0f / 512);
        const float rA = rsqrtf(fmaxf(tot[5] * (1.0f / 512) - mA * mA, 0.f) + 1e-5f);
        const float rB = rsqrtf(fmaxf(tot[7] * (1.0f / 512) - mB * mB, 0.f) + 1e-5f);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          F[0][b] = rA * (tot[b] - mA * my_cs[b]) + my_bw[b];
          F[1][b] = rB * (tot[2 + b] - mB * my_cs[b]) + my_bw[b];
        }
      }
      const int64_t f = (tile * 16) / p.rows_per_frame;  // rows_per_frame % 16 == 0: one frame per tile
      const float sig = p.sigma[f * p.sigma_stride];
      const float sig_o = p.sigma_other ? p.sigma_other[f * p.sigma_other_stride] : 0.f;
      const float c_skip = sd * sd / (sig * sig + sd * sd);
      const float c_out = sig * sd / sqrtf(sig * sig + sd * sd);
      const float sig_next = MODE == 1 ? sig_o : sig;  // sigma at which the NEXT evaluation runs
      const float c_in = 1.0f / sqrtf(sd * sd + sig_next * sig_next);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int64_t row = r == 0 ? rowA : rowB;
        float xs[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int ch = q * 8 + 2 * t + b;
          const bool ok = ch < p.C;
          const int64_t o = row * p.C + ch;
          const float x = xin[r][b];
          float xo;
          if (MODE == 3) {
            xo = x * sig;   // x_0 = latents * t_0
          } else if (MODE == 4) {
            xo = x;         // plain forward(): only the next projection h = proj_in(c_in x) is wanted
          } else {
            const float D = c_skip * x + c_out * F[r][b];
            if (MODE == 0) {
              xo = D;
            } else if (MODE == 1) {
              const float d = (x - D) / sig;
              xo = x + (sig_o - sig) * d;
              if (ok) p.d_buf[o] = d;
            } else {
              const float dp = (x - D) / sig;
              xo = xbase[r][b] + (sig - sig_o) * (0.5f * dcur[r][b] + 0.5f * dp);
            }
          }
          if (ok && p.x_out != nullptr) p.x_out[o] = xo;
          xs[b] = ok ? c_in * xo : 0.f;
        }
        uint32_t hi, lo;
        split_pair(xs[0], xs[1], hi, lo);
        my_xs[(r * 4 + q) * 32] = hi;
        my_xs[(8 + r * 4 + q) * 32] = lo;
      }
    }
    group_barrier(grp);  // c_in x_next is in shared memory; the partials may be overwritten by the next tile
    if (!need_next) continue;  // uniform over the CTA
    // ---- phase 3: h_next[:, 128 q : 128 q + 128] ----
    uint32_t xhi[2][4], xlo[2][4];   // [k16 step s][a0 .. a3]: channel tiles 2s (rows A, B), 2s + 1 (rows A, B)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      xhi[s][0] = my_xs[(0 + 2 * s) * 32];
      xhi[s][1] = my_xs[(4 + 2 * s) * 32];
      xhi[s][2] = my_xs[(0 + 2 * s + 1) * 32];
      xhi[s][3] = my_xs[(4 + 2 * s + 1) * 32];
      xlo[s][0] = my_xs[(8 + 2 * s) * 32];
      xlo[s][1] = my_xs[(12 + 2 * s) * 32];
      xlo[s][2] = my_xs[(8 + 2 * s + 1) * 32];
      xlo[s][3] = my_xs[(12 + 2 * s + 1) * 32];
    }
    float4* oA = reinterpret_cast<float4*>(p.h_next + rowA * 512) + t;
    float4* oB = reinterpret_cast<float4*>(p.h_next + rowB * 512) + t;
#pragma unroll 2
    for (int pp = 0; pp < 8; pp += 2) {
      // two column pairs at once: four independent accumulators
      const int pair = q * 8 + pp;
      float c[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a) c[a][0] = c[a][1] = c[a][2] = c[a][3] = 0.f;
      const uint4* wp = reinterpret_cast<const uint4*>(s_win + 16384) + pair * 128 + lane;
      uint4 whi[4], wlo[4];   // [pair pp: ab 0, 1 | pair pp + 1: ab 0, 1]
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        whi[a] = wp[a * 64];
        wlo[a] = wp[a * 64 + 32];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) mma_bf16(c[a], xlo[0], whi[a].x, whi[a].y);
#pragma unroll
      for (int a = 0; a < 4; ++a) mma_bf16(c[a], xhi[0], wlo[a].x, wlo[a].y);
#pragma unroll
      for (int a = 0; a < 4; ++a) mma_bf16(c[a], xhi[0], whi[a].x, whi[a].y);
#pragma unroll
      for (int a = 0; a < 4; ++a) mma_bf16(c[a], xlo[1], whi[a].z, whi[a].w);
#pragma unroll
      for (int a = 0; a < 4; ++a) mma_bf16(c[a], xhi[1], wlo[a].z, wlo[a].w);
#pragma unroll
      for (int a = 0; a < 4; ++a) mma_bf16(c[a], xhi[1], whi[a].z, whi[a].w);
      oA[pair * 4] = make_float4(c[0][0], c[0][1], c[1][0], c[1][1]);
      oB[pair * 4] = make_float4(c[0][2], c[0][3], c[1][2], c[1][3]);
      oA[pair * 4 + 4] = make_float4(c[2][0], c[2][1], c[3][0], c[3][1]);
      oB[pair * 4 + 4] = make_float4(c[2][2], c[2][3], c[3][2], c[3][3]);
    }
  }
}

template <int MODE, bool NEXT, int GROUPS>
static int launch_boundary_g(const BoundaryParams& p, int64_t T, cudaStream_t stream) {
  const int smem = BD_SMEM_WORDS * sizeof(uint32_t);
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(boundary_kernel<GROUPS, MODE, NEXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int64_t blocks = T / 16;  // one tile per CTA until every SM has one, then up to GROUPS tiles in flight per SM
  const int64_t cap = device_sm_count();
  if (blocks > cap) blocks = cap;
  RALD_CHECK_CUDA(launch_pdl(boundary_kernel<GROUPS, MODE, NEXT>, dim3((unsigned)blocks), dim3(GROUPS * 128), smem, stream, p));
  return 0;
}

template <int MODE, bool NEXT>
static int launch_boundary_as(const BoundaryParams& p, int64_t T, cudaStream_t stream) {
  // four groups of four warps (128 registers per thread); three groups (168 registers, no fewer spills: the
  // compiler hoists the weight fragments of all four column blocks) measured 42.3 vs 39.6 us per 32 768 rows
  return launch_boundary_g<MODE, NEXT, BD_MAX_GROUPS>(p, T, stream);
}

// the update mode and the presence of the next projection are compile-time properties of the kernel
static int launch_boundary(const BoundaryParams& p, int64_t T, cudaStream_t stream) {
  const bool next = p.h_next != nullptr;
  switch (p.mode) {
    case 0: return next ? launch_boundary_as<0, true>(p, T, stream) : launch_boundary_as<0, false>(p, T, stream);
    case 1: return next ? launch_boundary_as<1, true>(p, T, stream) : launch_boundary_as<1, false>(p, T, stream);
    case 2: return next ? launch_boundary_as<2, true>(p, T, stream) : launch_boundary_as<2, false>(p, T, stream);
    case 3: return next ? launch_boundary_as<3, true>(p, T, stream) : launch_boundary_as<3, false>(p, T, stream);
    default: return next ? launch_boundary_as<4, true>(p, T, stream) : launch_boundary_as<4, false>(p, T, stream);
  }
}

int dit_boundary_pack(const float* ln_w, const float* ln_b, const float* w_out_t, const float* w_in_t, int C,
                      void* pack, cudaStream_t stream) {
  RALD_REQUIRE(C >= 1 && C <= 32, "dit_boundary_pack: channels=%d must be in [1, 32]", C);
  RALD_REQUIRE(pack != nullptr && (reinterpret_cast<uintptr_t>(pack) & 15) == 0, "dit_boundary_pack: pack must be 16-byte aligned");
  boundary_pack_kernel<<<1, 512, 0, stream>>>(ln_w, ln_b, w_out_t, w_in_t, C, static_cast<uint32_t*>(pack));
  RALD_LAUNCHED();
  return 0;
}

int64_t dit_boundary_pack_bytes() { return (int64_t)BD_PACK_WORDS * 4; }

// scratch pack of a stream for callers that bring raw weights (one per stream: launches of one stream are ordered)
static void* boundary_scratch_pack(cudaStream_t stream) {
  static std::mutex mu;
  static std::unordered_map<cudaStream_t, void*> packs;
  std::lock_guard<std::mutex> lock(mu);
  auto it = packs.find(stream);
  if (it != packs.end()) return it->second;
  void* ptr = nullptr;
  if (cudaMalloc(&ptr, BD_PACK_WORDS * 4) != cudaSuccess) return nullptr;
  packs.emplace(stream, ptr);
  return ptr;
}

int dit_boundary(const float* h, const float* ln_w, const float* ln_b, const float* w_out_t, const float* w_in_t,
                 const float* x_in, const float* x_base, float* d_buf, float* x_out, float* h_next,
                 const float* sigma, int64_t sigma_stride, const float* sigma_other, int64_t sigma_other_stride,
                 int mode, int rows_per_frame, int C, int64_t T, int dim, float sigma_data, cudaStream_t stream,
                 const void* pack) {
  RALD_REQUIRE(dim == 512, "dit_boundary: dim=%d unsupported (512 only)", dim);
  RALD_REQUIRE(C >= 1 && C <= 32, "dit_boundary: channels=%d must be in [1, 32]", C);
  RALD_REQUIRE(mode >= 0 && mode <= 4, "dit_boundary: mode %d", mode);
  RALD_REQUIRE(mode >= 3 || h != nullptr, "dit_boundary: h missing");
  RALD_REQUIRE(rows_per_frame > 0 && rows_per_frame % 16 == 0 && T % 16 == 0,
               "dit_boundary: rows_per_frame=%d / T=%lld must be multiples of the 16-row tile", rows_per_frame, (long long)T);
  BoundaryParams p;
  p.pack_fresh = 0;
  if (pack == nullptr) {
    // raw weights: split + pack them in front of the launch (the runtimes pack once per weight version instead)
    void* scratch = boundary_scratch_pack(stream);
    RALD_REQUIRE(scratch != nullptr, "dit_boundary: no memory for the weight pack");
    RALD_TRY(dit_boundary_pack(ln_w, ln_b, w_out_t, w_in_t, C, scratch, stream));
    pack = scratch;
    p.pack_fresh = 1;
  }
  p.pack = static_cast<const uint32_t*>(pack);
  p.h = h; p.ln_w = ln_w; p.ln_b = ln_b; p.w_out_t = w_out_t; p.w_in_t = w_in_t; p.x_in = x_in; p.x_base = x_base;
  p.d_buf = d_buf; p.x_out = x_out; p.h_next = h_next; p.sigma = sigma; p.sigma_other = sigma_other;
  p.sigma_stride = sigma_stride; p.sigma_other_stride = sigma_other_stride; p.mode = mode;
  p.rows_per_frame = rows_per_frame; p.C = C; p.T = T; p.sigma_data = sigma_data;
  ProfScope prof(FAM_BOUNDARY, stream, (double)T * (mode < 3 ? 2048.0 : 0.0) + (h_next ? (double)T * 2048.0 : 0.0) +
                                          (double)T * C * 16.0);
  RALD_TRY(launch_boundary(p, T, stream));
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// radar tokens: tok[b, (r a e), :] = W x[b, r, a, e, :] + bias + r_emb[r] + a_emb[a] + e_emb[e]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
radar_tokens_kernel(const float* __restrict__ feat, int cz, int nr, int na, int ne, const float* __restrict__ w,
                    const float* __restrict__ b, const float* __restrict__ r_emb, const float* __restrict__ a_emb,
                    const float* __restrict__ e_emb, int dim, float* __restrict__ tok_f32,
                    __nv_bfloat16* __restrict__ tok_bf16) {
  const int64_t t = blockIdx.x;  // global token index over B * nr * na * ne
  const int e = (int)(t % ne);
  const int a = (int)((t / ne) % na);
  const int r = (int)((t / ((int64_t)ne * na)) % nr);
  const float* x = feat + t * cz;
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < cz; ++c) acc = fmaf(w[(int64_t)j * cz + c], x[c], acc);
    acc += b[j];
    acc = ((acc + r_emb[(int64_t)r * dim + j]) + a_emb[(int64_t)a * dim + j]) + e_emb[(int64_t)e * dim + j];
    if (tok_f32) tok_f32[t * dim + j] = acc;
    if (tok_bf16) tok_bf16[t * dim + j] = __float2bfloat16_rn(acc);
  }
}

int radar_tokens(const float* feat, int B, int nr, int na, int ne, int cz, const float* w, const float* b,
                 const float* r_emb, const float* a_emb, const float* e_emb, int dim, float* tok_f32, void* tok_bf16,
                 cudaStream_t stream) {
  const int64_t ntok = (int64_t)B * nr * na * ne;
  RALD_REQUIRE(ntok > 0 && ntok < (1ll << 31), "radar_tokens: bad token count");
  radar_tokens_kernel<<<(unsigned)ntok, 128, 0, stream>>>(feat, cz, nr, na, ne, w, b, r_emb, a_emb, e_emb, dim,
                                                          tok_f32, reinterpret_cast<__nv_bfloat16*>(tok_bf16));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald
