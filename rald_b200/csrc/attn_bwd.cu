// Backward of softmax(Q K^T * scale) V for head_dim 64 on tcgen05 (training step of the denoiser, SURVEY.md §8(f)
// row 3; forward = attn.cu, reference = CrossAttention.forward, model/models_radar_generation.py:58-76 under autograd).
//
// With P = softmax(S), S = scale Q K^T, O = P V and D_i = sum_d dO_id O_id:
//     dV = P^T dO            dP = dO V^T            dS = scale * P o (dP - D)            dQ = dS K            dK = dS^T Q
// P is never stored by the forward pass: it is recomputed from Q, K and the forward's row statistics
// lse2_i = m_i + log2(l_i) (log2 units, scale folded in), p_ij = 2^(s_ij scale log2e - lse2_i).
//
// Conditioning (measured, tools/probe_train_precision.py): dP_ij and D_i both contain dO_i . vbar, vbar = the component
// all values of a (frame, head) have in common, and only their DIFFERENCE enters dS. With V rounded to bf16 and
// D = rowsum(dO o O) taken from the bf16 forward output (the usual flash-attention form) the two copies of that term
// carry independent rounding errors of size 2^-9 |dO| |vbar|: on a randomly initialised deep network (nearly collinear
// tokens, |vbar| >> spread of V) that put 5 - 15 % of noise on the to_q / to_k weight gradients of the last blocks.
// Hence (a) the caller passes V CENTRED per (frame, head) — V - mean_j V_j, rounded to bf16 afterwards
// (rald_center_cast_f16_bf16): dS is invariant under a shift of all values, dV does not read V — and (b) D_i is not
// taken from O but computed as sum_j p_ij dP_ij from the very dP and P the kernel uses (first pass of MODE_DQ), so that
// every row of dS sums to zero to fp32 accuracy. With both, those gradients are within 1e-3 of fp32 autograd in the probe.
//
// Launches of ONE kernel template, each deterministic (no atomics):
//   MODE_DKV  CTA = (frame, head, 128-key tile): K_j, V_j resident in shared memory, the query tiles (Q_i, dO_i) stream
//             through a 2-stage TMA ring. Per query tile:  S^T = K_j Q_i^T and dP^T = V_j dO_i^T into TMEM (keys on the
//             TMEM lanes), four warps (thread <-> key row) turn them into P^T and dS^T (bf16, written back to TMEM as
//             A operands), then dV += P^T dO_i and dK += dS^T Q_i with the SAME shared-memory tiles of dO_i / Q_i read
//             as MN-major B operands. The per-query statistics are per COLUMN here and come from shared memory.
//   MODE_DQ   (runs first) CTA = (frame, head, 128-query tile): Q_i, dO_i resident, (K_j, V_j) stream TWICE. Pass 1:
//             S = Q_i K_j^T, dP = dO_i V_j^T, D_i = sum_j p_ij dP_ij (per lane, written to global for MODE_DKV). Pass 2:
//             S, dP again, dS (bf16 pair in TMEM), dQ += dS K_j (K_j as MN-major B operand).
// S and dP are recomputed in every pass (9 tile products instead of the minimal 5): the alternative is a dQ
// accumulated across CTAs with atomics or a TMA reduce-add, whose summation order would vary from run to run.
// In MODE_DQ dS is handed to the tensor cores as a SPLIT pair dS = hi + lo of bf16 values (16 mantissa bits, two products
// per accumulator): every row of dS sums to zero, so dQ_i = sum_j dS_ij K_j cancels whatever the keys have in common, and
// independently rounded entries leave |mean key| x sum_j eps_ij behind (unit test: dQ 3.0e-3 -> 2.4e-3 rel-L2). dK has no
// such cancellation (the columns of dS do not sum to zero) and takes the plain bf16 dS^T.
// TMEM columns: S [0,128), dP [128,256), then MODE_DQ: dS_hi [256,320), dS_lo [320,384), dQ [384,448);
// MODE_DKV: P^T [256,320), dS^T [320,384), dV [384,448), dK [448,512). Nothing is written over S / dP: the two compute
// warps of a lane quarter work on different column halves at their own pace.
#include "../../include/rald_b200.h"

#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int AB_THREADS = 320;              // warp 0: TMA, warp 1: MMA issue, warps 2..9: compute — two per TMEM lane
                                             // quarter (a scheduler with ONE such warp cannot hide its ex2 / TMEM latencies),
                                             // each taking half of the tile's columns
constexpr int AB_TILE = 128 * 64 * 2;        // one [128 x 64] 16-bit tile, 128-byte swizzled rows
constexpr int AB_MODE_DQ = 0, AB_MODE_DKV = 1;

struct AttnBwdParams {
  int Sq, Skv, frames, heads;
  int r_tiles;           // resident tiles per (frame, head): q tiles (DQ) or key tiles (DKV)
  int x_tiles;           // streamed tiles per (frame, head)
  int r_frame_rows;      // rows per frame of the resident operand (Sq or Skv)
  int x_frame_rows;      // rows per frame of the streamed operand
  int x_rows;            // rows per streamed tile (128, or 64 for the 64-token context in MODE_DQ)
  float scale_log2, scale;
  const float* lse2;     // [frames][heads][Sq]
  float* dsum;           // [frames][heads][Sq]: D_i, written by MODE_DQ (pass 1), read by MODE_DKV
  uint16_t* out1;        // DQ: dQ ; DKV: dV
  int64_t ld1;
  uint16_t* out2;        // DKV: dK
  int64_t ld2;
  uint32_t idesc_s, idesc_dp, idesc_acc;
};

template <int MODE>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
                const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sR1 = smem;                    // DQ: Q_i   DKV: K_j
  uint8_t* sR2 = sR1 + AB_TILE;           // DQ: dO_i  DKV: V_j
  uint8_t* sX1 = sR2 + AB_TILE;           // [2]  DQ: K_j   DKV: Q_i
  uint8_t* sX2 = sX1 + 2 * AB_TILE;       // [2]  DQ: V_j   DKV: dO_i
  float* s_lse = reinterpret_cast<float*>(sX2 + 2 * AB_TILE);   // [2][128]
  float* s_ds = s_lse + 2 * 128;                                 // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_ds + 2 * 128);
  uint64_t* r_full = bars + 0;
  uint64_t* x_full = bars + 1;    // [2]
  uint64_t* x_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmR2);
    tma_prefetch_desc(&tmX1);
    tma_prefetch_desc(&tmX2);
    mbar_init(r_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, 8);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int item = blockIdx.x;
  const int fh = item / p.r_tiles;
  const int rt = item - fh * p.r_tiles;
  const int frame = fh / p.heads, head = fh - frame * p.heads;
  const int nx = p.x_tiles;
  const int iters = MODE == AB_MODE_DQ ? 2 * nx : nx;   // MODE_DQ streams the keys twice (D pass, then dS / dQ pass)
  constexpr uint32_t COL_S = 0, COL_DP = 128, COL_P = 256, COL_DS = MODE == AB_MODE_DKV ? 320 : 256, COL_LO = 320,
                     COL_A1 = 384, COL_A2 = 448;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      const int r_row = frame * p.r_frame_rows + rt * 128;
      mbar_arrive_expect_tx(r_full, 2 * AB_TILE);
      tma_load_2d(sR1, &tmR1, r_full, head * 64, r_row);
      tma_load_2d(sR2, &tmR2, r_full, head * 64, r_row);
      const uint32_t x_bytes = 2u * (uint32_t)p.x_rows * 128u;
      for (int it = 0; it < iters; ++it) {
        const int st = it & 1;
        const int s = it >= nx ? it - nx : it;
        mbar_wait(&x_empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&x_full[st], x_bytes);
        const int x_row = frame * p.x_frame_rows + s * p.x_rows;
        tma_load_2d(sX1 + st * AB_TILE, &tmX1, &x_full[st], head * 64, x_row);
        tma_load_2d(sX2 + st * AB_TILE, &tmX2, &x_full[st], head * 64, x_row);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      mbar_wait(r_full, 0);
      tc_fence_after();
      const uint64_t r1_desc = make_sdesc_sw128(smem_u32(sR1), 16, 1024);
      const uint64_t r2_desc = make_sdesc_sw128(smem_u32(sR2), 16, 1024);
      const int acc_k = (MODE == AB_MODE_DQ ? p.x_rows : 128) / 16;   // 16-deep steps of the accumulating products
      for (int it = 0; it < iters; ++it) {
        const int st = it & 1;
        const int s = it >= nx ? it - nx : it;
        const bool acc_pass = MODE != AB_MODE_DQ || it >= nx;
        mbar_wait(&x_full[st], (it >> 1) & 1);
        tc_fence_after();
        const uint32_t x1 = smem_u32(sX1 + st * AB_TILE), x2 = smem_u32(sX2 + st * AB_TILE);
        const uint64_t x1_desc = make_sdesc_sw128(x1, 16, 1024), x2_desc = make_sdesc_sw128(x2, 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_f16_ss(tmem_base + COL_S, r1_desc + 2 * k, x1_desc + 2 * k, p.idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_f16_ss(tmem_base + COL_DP, r2_desc + 2 * k, x2_desc + 2 * k, p.idesc_dp, k != 0);
        tc_commit(s_full);
        mbar_wait(p_ready, it & 1);
        tc_fence_after();
        if (!acc_pass) {           // D pass: nothing to accumulate, the tiles are free once S and dP have been read
          tc_commit(&x_empty[st]);
          continue;
        }
        // the streamed tiles again, as MN-major B operands ([k index][64 values of d]): 8-row groups 1024 B apart
        const uint64_t x1_mn = make_sdesc_sw128(x1, 1024, 1024), x2_mn = make_sdesc_sw128(x2, 1024, 1024);
        if (MODE == AB_MODE_DKV) {
          for (int k = 0; k < acc_k; ++k)   // dV += P^T dO_i
            mma_f16_ts(tmem_base + COL_A1, tmem_base + COL_P + 8 * k, x2_mn + 128 * k, p.idesc_acc, (s | k) != 0);
          for (int k = 0; k < acc_k; ++k)   // dK += dS^T Q_i
            mma_f16_ts(tmem_base + COL_A2, tmem_base + COL_DS + 8 * k, x1_mn + 128 * k, p.idesc_acc, (s | k) != 0);
        } else {
          for (int k = 0; k < acc_k; ++k)   // dQ += dS K_j   (hi, then lo)
            mma_f16_ts(tmem_base + COL_A1, tmem_base + COL_DS + 8 * k, x1_mn + 128 * k, p.idesc_acc, (s | k) != 0);
          for (int k = 0; k < acc_k; ++k)
            mma_f16_ts(tmem_base + COL_A1, tmem_base + COL_LO + 8 * k, x1_mn + 128 * k, p.idesc_acc, 1);
        }
        tc_commit(&x_empty[st]);   // the streamed tiles may be replaced once these products have retired
      }
      tc_commit(acc_done);
    }
  } else {
    // ===================== P / dS computation + epilogue (warps 2..9) =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int ch = (warp - 2) >> 2;               // column half of the tile this warp works on
    const int row = q * 32 + lane;                // tile row = TMEM lane
    const uint32_t t_mine = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int64_t stat_base = ((int64_t)frame * p.heads + head) * p.Sq;
    float lse_r = 0.f, ds_r = 0.f;
    if (MODE == AB_MODE_DQ) lse_r = p.lse2[stat_base + rt * 128 + row];
    const int ncols = MODE == AB_MODE_DQ ? p.x_rows : 128;
    const int hc = ncols >> 1;                    // columns per warp (64, or 32 for a 64-token context in MODE_DQ)
    const int cbeg = ch * hc;
    for (int it = 0; it < iters; ++it) {
      const int s = it >= nx ? it - nx : it;
      const float* lse_c = s_lse + (s & 1) * 128;
      const float* ds_c = s_ds + (s & 1) * 128;
      if (MODE == AB_MODE_DQ && it < nx) {
        // ---- pass 1: D_i = sum_j p_ij dP_ij over all keys (four independent chains per 32 columns) ----
        mbar_wait(s_full, it & 1);
        tc_fence_after();
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 1
        for (int c = cbeg; c < cbeg + hc; c += 32) {
          uint32_t sv[32], dv[32];
          tmem_ld32(t_mine + COL_S + c, sv);
          tmem_ld32(t_mine + COL_DP + c, dv);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            a0 = fmaf(ex2_f32(fmaf(__uint_as_float(sv[j]), p.scale_log2, -lse_r)), __uint_as_float(dv[j]), a0);
            a1 = fmaf(ex2_f32(fmaf(__uint_as_float(sv[j + 1]), p.scale_log2, -lse_r)), __uint_as_float(dv[j + 1]), a1);
            a2 = fmaf(ex2_f32(fmaf(__uint_as_float(sv[j + 2]), p.scale_log2, -lse_r)), __uint_as_float(dv[j + 2]), a2);
            a3 = fmaf(ex2_f32(fmaf(__uint_as_float(sv[j + 3]), p.scale_log2, -lse_r)), __uint_as_float(dv[j + 3]), a3);
          }
        }
        ds_r += (a0 + a1) + (a2 + a3);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_ready);
        if (it == nx - 1) {
          // the two column halves of a row meet: D_i = half 0 + half 1 (fixed order), written once for MODE_DKV
          s_lse[ch * 128 + row] = ds_r;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          ds_r = s_lse[row] + s_lse[128 + row];
          if (ch == 0) p.dsum[stat_base + rt * 128 + row] = ds_r;
        }
        continue;
      }
      if (MODE == AB_MODE_DKV) {
        // statistics of the 128 queries of this tile = the COLUMNS of S^T: staged in shared memory, read as broadcasts
        if (ch == 0) {
          s_lse[(s & 1) * 128 + row] = p.lse2[stat_base + s * 128 + row];
          s_ds[(s & 1) * 128 + row] = p.dsum[stat_base + s * 128 + row];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(s_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = cbeg; c < cbeg + hc; c += 32) {
        uint32_t sv[32], dv[32];
        tmem_ld32(t_mine + COL_S + c, sv);
        tmem_ld32(t_mine + COL_DP + c, dv);
        tmem_ld_wait();
        uint32_t pp[16], dd[16], dl[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float l0, l1, d0, d1;
          if (MODE == AB_MODE_DKV) {
            l0 = lse_c[c + 2 * j]; l1 = lse_c[c + 2 * j + 1];
            d0 = ds_c[c + 2 * j]; d1 = ds_c[c + 2 * j + 1];
          } else {
            l0 = l1 = lse_r; d0 = d1 = ds_r;
          }
          const float p0 = ex2_f32(fmaf(__uint_as_float(sv[2 * j]), p.scale_log2, -l0));
          const float p1 = ex2_f32(fmaf(__uint_as_float(sv[2 * j + 1]), p.scale_log2, -l1));
          const float g0 = p0 * (__uint_as_float(dv[2 * j]) - d0) * p.scale;
          const float g1 = p1 * (__uint_as_float(dv[2 * j + 1]) - d1) * p.scale;
          pp[j] = pack_bf16x2(p0, p1);
          dd[j] = pack_bf16x2(g0, g1);
          if (MODE == AB_MODE_DQ)
            dl[j] = pack_bf16x2(g0 - __uint_as_float(dd[j] << 16), g1 - __uint_as_float(dd[j] & 0xffff0000u));
        }
        if (MODE == AB_MODE_DKV) tmem_st16(t_mine + COL_P + (c >> 1), pp);
        tmem_st16(t_mine + COL_DS + (c >> 1), dd);
        if (MODE == AB_MODE_DQ) tmem_st16(t_mine + COL_LO + (c >> 1), dl);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // ---- epilogue: accumulators -> bf16 rows of the gradient tensors (each warp: 32 of the 64 columns) ----
    mbar_wait(acc_done, 0);
    tc_fence_after();
    const int tile_row = rt * 128 + row;
    const bool valid = tile_row < p.r_frame_rows;
    const int64_t grow = (int64_t)frame * p.r_frame_rows + tile_row;
    auto store32 = [&](uint32_t col, uint16_t* out, int64_t ld) {
      uint32_t v[32];
      tmem_ld32(t_mine + col + 32 * ch, v);
      tmem_ld_wait();
      if (valid) {
        uint4* dst = reinterpret_cast<uint4*>(out + grow * ld + head * 64 + 32 * ch);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1])),
                              pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                              pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                              pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
      }
    };
    store32(COL_A1, p.out1, p.ld1);
    if (MODE == AB_MODE_DKV) store32(COL_A2, p.out2, p.ld2);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// lse2[f][h][r] = m + log2(l) from the forward's statistics [frames*Sq][heads][2]
__global__ void __launch_bounds__(256)
attn_bwd_prep_kernel(const float* __restrict__ stats, int64_t rows, int Sq, int heads, float* __restrict__ lse2) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * heads) return;
  const int64_t row = i / heads;
  const int head = (int)(i - row * heads);
  const int64_t frame = row / Sq;
  const int r = (int)(row - frame * Sq);
  lse2[(frame * heads + head) * Sq + r] = stats[i * 2] + log2f(stats[i * 2 + 1]);
}

int attn_d64_bwd(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                 const void* dO, int64_t lddo, const float* stats, float* lse2, float* dsum,
                 void* dQ, int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, int frames, int heads, int Sq,
                 int Skv, float scale, cudaStream_t stream) {
  RALD_REQUIRE(frames > 0 && heads > 0 && heads <= 8, "attn_bwd: bad sizes (heads <= 8)");
  RALD_REQUIRE(Sq % 128 == 0, "attn_bwd: Sq=%d must be a multiple of 128", Sq);
  RALD_REQUIRE(Skv == 64 || Skv % 128 == 0, "attn_bwd: Skv=%d must be 64 or a multiple of 128", Skv);
  RALD_REQUIRE(lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0 && lddo % 8 == 0 && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0,
               "attn_bwd: row pitches must be multiples of 8 elements");
  const int64_t rows = (int64_t)frames * Sq;
  attn_bwd_prep_kernel<<<(unsigned)((rows * heads + 255) / 256), 256, 0, stream>>>(stats, rows, Sq, heads, lse2);
  RALD_LAUNCHED();

  const int q_tiles = Sq / 128;
  const int kv_tiles = (Skv + 127) / 128;
  const int nk = Skv < 128 ? Skv : 128;
  const uint64_t cols = (uint64_t)heads * 64;
  const int smem_bytes = 6 * AB_TILE + 4 * 128 * 4 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<AB_MODE_DQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes));
    RALD_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<AB_MODE_DKV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes));
    configured = true;
  }
  CUtensorMap tmQ, tmdO, tmK, tmV;
  RALD_TRY(make_tmap_2d_bf16(&tmQ, Q, (uint64_t)rows, cols, (uint64_t)ldq, 128));
  RALD_TRY(make_tmap_2d_bf16(&tmdO, dO, (uint64_t)rows, cols, (uint64_t)lddo, 128));
  AttnBwdParams p;
  p.Sq = Sq; p.Skv = Skv; p.frames = frames; p.heads = heads;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse2 = lse2;
  p.dsum = dsum;
  p.idesc_acc = make_idesc(FMT_BF16, 128, 64, 0, 1);
  {
    // ---- D and dQ: queries resident, keys streamed twice in tiles of nk rows ----
    RALD_TRY(make_tmap_2d_bf16(&tmK, K, (uint64_t)frames * Skv, cols, (uint64_t)ldk, (uint32_t)nk));
    RALD_TRY(make_tmap_2d_bf16(&tmV, V, (uint64_t)frames * Skv, cols, (uint64_t)ldv, (uint32_t)nk));
    p.r_tiles = q_tiles; p.x_tiles = kv_tiles;
    p.r_frame_rows = Sq; p.x_frame_rows = Skv; p.x_rows = nk;
    p.out1 = reinterpret_cast<uint16_t*>(dQ); p.ld1 = lddq;
    p.out2 = nullptr; p.ld2 = 0;
    p.idesc_s = make_idesc(FMT_BF16, 128, (uint32_t)nk, 0, 0);
    p.idesc_dp = make_idesc(FMT_BF16, 128, (uint32_t)nk, 0, 0);
    ProfScope prof(FAM_ATTN, stream, 10.0 * frames * heads * Sq * (double)Skv * 64);
    attn_bwd_kernel<AB_MODE_DQ><<<frames * heads * q_tiles, AB_THREADS, smem_bytes, stream>>>(tmQ, tmdO, tmK, tmV, p);
    RALD_LAUNCHED();
  }
  {
    // ---- dK, dV: keys resident (128-row boxes: rows past a 64-key context belong to the next frame or are zero-filled;
    // they only reach TMEM lanes whose results are not stored) ----
    RALD_TRY(make_tmap_2d_bf16(&tmK, K, (uint64_t)frames * Skv, cols, (uint64_t)ldk, 128));
    RALD_TRY(make_tmap_2d_bf16(&tmV, V, (uint64_t)frames * Skv, cols, (uint64_t)ldv, 128));
    p.r_tiles = kv_tiles; p.x_tiles = q_tiles;
    p.r_frame_rows = Skv; p.x_frame_rows = Sq; p.x_rows = 128;
    p.out1 = reinterpret_cast<uint16_t*>(dV); p.ld1 = lddv;
    p.out2 = reinterpret_cast<uint16_t*>(dK); p.ld2 = lddk;
    p.idesc_s = make_idesc(FMT_BF16, 128, 128, 0, 0);
    p.idesc_dp = make_idesc(FMT_BF16, 128, 128, 0, 0);
    ProfScope prof(FAM_ATTN, stream, 8.0 * frames * heads * Sq * (double)(kv_tiles * 128) * 64);
    attn_bwd_kernel<AB_MODE_DKV><<<frames * heads * kv_tiles, AB_THREADS, smem_bytes, stream>>>(tmK, tmV, tmQ, tmdO, p);
    RALD_LAUNCHED();
  }
  return 0;
}

}  // namespace rald

extern "C" int rald_attn_d64_stats(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                   void* O, int64_t ldo, int frames, int heads, int Sq, int Skv, float scale,
                                   float* stats, void* stream) {
  return rald::attn_d64_chunk(Q, ldq, K, ldk, V, ldv, O, ldo, frames, heads, Sq, Skv, Skv, stats, scale,
                              static_cast<cudaStream_t>(stream));
}

extern "C" int rald_attn_d64_bwd(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V_centred, int64_t ldv,
                                 const void* dO, int64_t lddo, const float* stats, float* lse2_ws, float* dsum_ws, void* dQ,
                                 int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, int frames, int heads, int Sq,
                                 int Skv, float scale, void* stream) {
  return rald::attn_d64_bwd(Q, ldq, K, ldk, V_centred, ldv, dO, lddo, stats, lse2_ws, dsum_ws, dQ, lddq, dK, lddk, dV, lddv,
                            frames, heads, Sq, Skv, scale, static_cast<cudaStream_t>(stream));
}
