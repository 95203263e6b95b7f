// Fused cross-attention sub-layer of the denoiser block for a short, per-sample-constant context
// (model/models_radar_generation.py:35-76 as called at :167, attn2 of BasicTransformerBlock):
//     h += to_out( softmax( to_q(xn) to_k(ctx)^T / sqrt(64) ) to_v(ctx) ) + b_out          per head, 8 x 64
// The conditioning tokens do not change over the 35 network evaluations of a sample, so the two 512 x 512 linears
// are folded ONCE per sample into the per-frame context operands (xattn_fold below, plain tcgen05 GEMMs):
//     K'[h][f][key][i] = c * sum_d K[f][key][h*64+d] * Wq[h*64+d][i]        c = log2(e) / sqrt(64)
//     VT[h][o][f][key] =     sum_d Wo[o][h*64+d]   * V[f][key][h*64+d]                  (stored as fp16)
// after which the whole sub-layer is two GEMM-shaped products around a softmax over each head's 64 keys:
//     S[row][(h,key)] = xn[row][:] . K'[h][f][key][:]         (M = 128 rows, N = 512 = 8 heads x 64 keys, K = 512)
//     P = softmax over each group of 64 columns               (registers <-> TMEM, fp16 probabilities)
//     h[row][o] += sum_(h,key) P[row][(h,key)] VT[h][o][f][key] + b[o]       (K = 512, N = 512; TMA reduce-add)
// One kernel replaces to_q GEMM -> attention -> to_out GEMM: the q and attention-output activations
// (4 x 1 KB per row) never exist and the tiny-context attention (5 % tensor utilisation on its own) disappears.
//
// Per 128-row tile, one CTA per SM, persistent:
//   warp 0 lane 0  TMA producer: xn k-blocks (ring A, 2 x 16 KB), K' / VT tiles (ring B, 4 x 32 KB)
//   warp 1 lane 0  tcgen05 issuer: S = 8 k-steps x 2 MMAs (N = 256) into all 512 TMEM columns;
//                  O in four quarters of 128 output columns, A operand = P read from TMEM (kind::f16, fp16 x fp16)
//   warps 2..9     softmax (warp set hs = 0: heads 0..3, hs = 1: heads 4..7; thread <-> row) then epilogue
// TMEM plan: S fills [0, 512). P of heads 0..3 overwrites [0, 128) (already consumed columns), P of heads 4..7
// overwrites [256, 384); the freed [128, 256) and [384, 512) are the two accumulator buffers of the O quarters.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int XA_BM = 128;
constexpr int XA_DIM = 512;          // model width = K of the S product = N of the O product
constexpr int XA_HK = 512;           // heads x keys = N of the S product = K of the O product
constexpr int XA_KEYS = 64;
constexpr int XA_THREADS = 320;
constexpr int XA_A_BYTES = XA_BM * 64 * 2;   // 16 KB
constexpr int XA_A_STAGES = 2;
constexpr int XA_B_TILE = 256 * 64 * 2;      // 32 KB: 256 K' rows x 64 k, or 2 x (128 VT rows x 64 k)
constexpr int XA_B_RING = 4 * XA_B_TILE;     // 128 KB of ring per CTA
constexpr int XA_STG_BYTES = 8 * 2 * 4096;   // two 32 x 32 fp32 staging tiles per epilogue warp
constexpr int XA_SMEM = XA_A_STAGES * XA_A_BYTES + XA_B_RING + XA_STG_BYTES + XA_DIM * 4 + 256;

struct XattnParams {
  const float* bias;   // [512] to_out bias
  int num_tiles;       // T / 128
  int tiles_per_frame; // n_latents / 128
  int frame0;          // first frame of this micro-batch inside the context operands
  int total_frames;    // frames the context operands were built for (row / column strides)
  unsigned long long* dbg;  // optional [tile < 4][16] %globaltimer stamps of CTA 0 (tools/gpu_time_xattn.py --phases)
};

static unsigned long long* g_xattn_dbg = nullptr;
#define XA_STAMP(slot_)                                                                    \
  do {                                                                                     \
    if (p.dbg != nullptr && blockIdx.x == 0 && it < 4) p.dbg[it * 16 + (slot_)] = global_timer_ns(); \
  } while (0)

// CG = 1: one CTA per 128-row tile. CG = 2: a CTA pair (cta_group::2) works on two neighbouring tiles of ONE frame with
// 256-row MMAs: each CTA stages its own xn rows but only HALF of every K' / VT tile (the pair's tensor cores share the
// halves), which halves the shared-memory fill + operand-read traffic per flop — the resource that paces the
// single-CTA form (fill + reads of a B tile that is used by one MMA group only run at the 128 B/clk crossbar limit).
// The leader (rank 0) issues all MMAs; softmax / epilogue warps of both CTAs work on their own 128 TMEM lanes.
template <int CG>
__global__ void __launch_bounds__(XA_THREADS, 1)
xattn_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   const XattnParams p) {
  constexpr int B_BYTES = XA_B_TILE / CG;          // this CTA's share of a B tile
  constexpr int B_STAGES = XA_B_RING / B_BYTES;    // 4 (CG = 1) or 8 (CG = 2)
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* ringA = smem;
  uint8_t* ringB = ringA + XA_A_STAGES * XA_A_BYTES;
  uint8_t* stg = ringB + XA_B_RING;
  float* s_bias = reinterpret_cast<float*>(stg + XA_STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + XA_DIM);
  uint64_t* a_full = bars;                     // [2]   (leader's copy is the one waited on when CG = 2)
  uint64_t* a_empty = a_full + XA_A_STAGES;    // [2]
  uint64_t* b_full = a_empty + XA_A_STAGES;    // [8]
  uint64_t* b_empty = b_full + 8;              // [8]
  uint64_t* s_full = b_empty + 8;              // [2] S complete (one per head set)
  uint64_t* p_full = s_full + 2;               // [2] P of head set 0 / 1 written by its softmax warps (of both CTAs)
  uint64_t* o_full = p_full + 2;               // [2] O quarter accumulated
  uint64_t* o_empty = o_full + 2;              // [2] O quarter drained by the epilogue warps (of both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int unit_tiles = p.num_tiles / CG;     // tile pairs when CG = 2 (tiles 2u, 2u + 1: same frame)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < XA_A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4 * CG);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 8 * CG);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  for (int i = threadIdx.x; i < XA_DIM; i += XA_THREADS) s_bias[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // (Prefetching the context boxes of this CTA's tiles into L2 here, before the grid dependency resolves, was measured
  // to HURT: the 160 prefetches per tile compete with the first tile's demand loads — 73 vs 50 us per launch in the
  // sampling loop.)
  pdl_wait();               // xn comes from the preceding LayerNorm kernel, h from the GEMM before it
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer (every CTA: its xn rows, its share of the B tiles) =====================
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      for (int ut = unit; ut < unit_tiles; ut += num_units) {
        const int tile = ut * CG + (int)rank;
        const int f = p.frame0 + tile / p.tiles_per_frame;
        // S phase: per k-step one xn tile and the two 256-column halves of K' (4 heads x 64 keys each). (Running the
        // two head sets one after the other, so that the first softmax overlaps the second set's MMAs, was measured
        // SLOWER: 60.5 vs 56.2 us per launch at 64 frames, and that order streams xn twice.)
        for (int kb = 0; kb < XA_DIM / 64; ++kb) {
          mbar_wait(&a_empty[sa], pha ^ 1);
          if (CG == 2) {
            // every byte of the pair is credited to the LEADER's full barrier, on which only the leader arrives
            if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2 * XA_A_BYTES);
            tma_load_2d_pair(ringA + sa * XA_A_BYTES, &tmX, &a_full[sa], kb * 64, tile * XA_BM);
          } else {
            mbar_arrive_expect_tx(&a_full[sa], XA_A_BYTES);
            tma_load_2d(ringA + sa * XA_A_BYTES, &tmX, &a_full[sa], kb * 64, tile * XA_BM);
          }
          if (++sa == XA_A_STAGES) { sa = 0; pha ^= 1; }
          for (int j = 0; j < 2; ++j) {
            mbar_wait(&b_empty[sb], phb ^ 1);
            if (CG == 2) {
              // rows [128 rank, 128 rank + 128) of the 256-row K' tile = heads 4 j + 2 rank + {0, 1}
              if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], XA_B_TILE);
              for (int i = 0; i < 2; ++i)
                tma_load_2d_pair(ringB + sb * B_BYTES + i * (XA_KEYS * 128), &tmK, &b_full[sb], kb * 64,
                                 ((4 * j + 2 * (int)rank + i) * p.total_frames + f) * XA_KEYS);
            } else {
              mbar_arrive_expect_tx(&b_full[sb], XA_B_TILE);
              for (int i = 0; i < 4; ++i)
                tma_load_2d(ringB + sb * B_BYTES + i * (XA_KEYS * 128), &tmK, &b_full[sb], kb * 64,
                            ((4 * j + i) * p.total_frames + f) * XA_KEYS);
            }
            if (++sb == B_STAGES) { sb = 0; phb ^= 1; }
          }
        }
        // O phase: per output quarter, four slots of two VT k-blocks (= two heads) each; with CG = 2 this CTA stages
        // output rows [64 rank, 64 rank + 64) of each 128-row quarter tile
        for (int q = 0; q < 4; ++q) {
          for (int kp = 0; kp < 4; ++kp) {
            mbar_wait(&b_empty[sb], phb ^ 1);
            if (CG == 2) {
              if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], XA_B_TILE);
              for (int sub = 0; sub < 2; ++sub)
                tma_load_2d_pair(ringB + sb * B_BYTES + sub * (B_BYTES / 2), &tmV, &b_full[sb], f * XA_KEYS,
                                 (2 * kp + sub) * XA_DIM + q * 128 + (int)rank * 64);
            } else {
              mbar_arrive_expect_tx(&b_full[sb], XA_B_TILE);
              for (int sub = 0; sub < 2; ++sub)
                tma_load_2d(ringB + sb * B_BYTES + sub * (B_BYTES / 2), &tmV, &b_full[sb], f * XA_KEYS,
                            (2 * kp + sub) * XA_DIM + q * 128);
            }
            if (++sb == B_STAGES) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader CTA only when CG = 2) =====================
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc_s = make_idesc(FMT_BF16, XA_BM * CG, 256, 0, 0);
      constexpr uint32_t idesc_o = make_idesc(FMT_F16, XA_BM * CG, 128, 0, 0);   // P fp16 (TMEM) x VT fp16 (smem)
      auto commit = [&](uint64_t* bar) {
        if (CG == 2) tc_commit_pair(bar);
        else tc_commit(bar);
      };
      auto wait_peer = [&](uint64_t* bar, uint32_t parity) {   // barriers the peer's warps arrive on remotely
        if (CG == 2) mbar_wait_cluster(bar, parity);
        else mbar_wait(bar, parity);
      };
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int it = 0;
      for (int ut = unit; ut < unit_tiles; ut += num_units, ++it) {
        // the previous tile's last two O quarters (one per buffer) have been drained; since those were committed after
        // every earlier MMA, all reads of the previous P have retired too: S may overwrite the whole TMEM
        const uint32_t u0 = 2u * (uint32_t)it;   // use index of each O buffer at this tile's first quarter
        wait_peer(&o_empty[0], (u0 & 1u) ^ 1u);
        wait_peer(&o_empty[1], (u0 & 1u) ^ 1u);
        tc_fence_after();
        XA_STAMP(0);
        for (int kb = 0; kb < XA_DIM / 64; ++kb) {
          mbar_wait(&a_full[sa], pha);
          const uint64_t a_desc = make_sdesc_sw128(smem_u32(ringA + sa * XA_A_BYTES), 16, 1024);
          for (int j = 0; j < 2; ++j) {
            mbar_wait(&b_full[sb], phb);
            tc_fence_after();
            const uint64_t b_desc = make_sdesc_sw128(smem_u32(ringB + sb * B_BYTES), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CG == 2) mma_f16_ss_pair(tmem_base + 256 * j, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (kb | k) != 0 ? 1u : 0u);
              else mma_f16_ss(tmem_base + 256 * j, a_desc + 2 * k, b_desc + 2 * k, idesc_s, (kb | k) != 0 ? 1u : 0u);
            }
            commit(&b_empty[sb]);
            if (++sb == B_STAGES) { sb = 0; phb ^= 1; }
          }
          commit(&a_empty[sa]);
          if (++sa == XA_A_STAGES) { sa = 0; pha ^= 1; }
        }
        commit(&s_full[0]);
        commit(&s_full[1]);
        XA_STAMP(1);
        for (int q = 0; q < 4; ++q) {
          const int buf = q & 1;
          const uint32_t u = u0 + (uint32_t)(q >> 1);
          wait_peer(&o_empty[buf], (u & 1u) ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (buf ? 384u : 128u);
          for (int kp = 0; kp < 4; ++kp) {
            // heads 0..3 read P of set 0, heads 4..7 of set 1 (written while the first MMAs of this quarter run)
            if (q == 0 && (kp == 0 || kp == 2)) {
              wait_peer(&p_full[kp >> 1], (uint32_t)it & 1u);
              XA_STAMP(2 + (kp >> 1));
            }
            mbar_wait(&b_full[sb], phb);
            tc_fence_after();
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const int kb = 2 * kp + sub;   // = head
              const uint32_t p_col = tmem_base + (kb < 4 ? 32u * kb : 256u + 32u * (kb - 4));
              const uint64_t b_desc = make_sdesc_sw128(smem_u32(ringB + sb * B_BYTES + sub * (B_BYTES / 2)), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (CG == 2) mma_f16_ts_pair(d_tmem, p_col + 8 * k, b_desc + 2 * k, idesc_o, (kb | k) != 0 ? 1u : 0u);
                else mma_f16_ts(d_tmem, p_col + 8 * k, b_desc + 2 * k, idesc_o, (kb | k) != 0 ? 1u : 0u);
              }
            }
            commit(&b_empty[sb]);
            if (++sb == B_STAGES) { sb = 0; phb ^= 1; }
          }
          commit(&o_full[buf]);
          XA_STAMP(4 + q);
        }
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..9 of every CTA, on its own tile) =====================
    const int qd = warp & 3;             // TMEM lane quarter
    const int ew = warp - 2;             // 0..7
    const int hs = ew >> 2;              // head set: 0 -> heads 0..3, 1 -> heads 4..7
    const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
    auto arrive_mma = [&](uint64_t* bar) {   // barriers the (leader's) MMA thread waits on
      if (CG == 2) mbar_arrive_leader(bar);
      else mbar_arrive(bar);
    };
    int it = 0;
    for (int ut = unit; ut < unit_tiles; ut += num_units, ++it) {
      const int tile = ut * CG + (int)rank;
      mbar_wait(&s_full[hs], (uint32_t)it & 1u);
      tc_fence_after();
      if (warp == 2 && lane == 0) XA_STAMP(8);
      const uint32_t t_set = tmem_base + lane_off + 256u * hs;   // S columns of this head set; P goes to its start
      // One head (64 score columns) at a time. The next head's columns are fetched while this one is processed, in two
      // halves (the second half is only issued once this head's first 32 scores are dead), which keeps the peak at ~130
      // live registers: the 168 available with 10 warps per CTA then hold the kernel without spills (round 1 prefetched
      // the whole next head up front and spilled 350 - 540 bytes per thread).
      auto head_probs = [&](const uint32_t (&v)[32], const uint32_t (&w)[32], uint32_t (&nv)[32], uint32_t (&nw)[32],
                            int g, bool more) {
        if (more) tmem_ld32(t_set + 64 * (g + 1), nv);
        float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(w[0]),
              m3 = __uint_as_float(w[1]);
#pragma unroll
        for (int j = 2; j < 32; j += 2) {
          m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
          m2 = fmaxf(m2, __uint_as_float(w[j])); m3 = fmaxf(m3, __uint_as_float(w[j + 1]));
        }
        const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        // scores are already in log2 units (c folded into K'): p = 2^(s - max) <= 1, packed fp16 pairs
        uint32_t lo[16], hi[16];
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          lo[j] = pack_f16x2(ex2_f32(__uint_as_float(v[2 * j]) - mx), ex2_f32(__uint_as_float(v[2 * j + 1]) - mx));
          const float2 a = unpack_f16x2(lo[j]);
          s0 += a.x; s1 += a.y;
        }
        if (more) tmem_ld32(t_set + 64 * (g + 1) + 32, nw);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          hi[j] = pack_f16x2(ex2_f32(__uint_as_float(w[2 * j]) - mx), ex2_f32(__uint_as_float(w[2 * j + 1]) - mx));
          const float2 b = unpack_f16x2(hi[j]);
          s2 += b.x; s3 += b.y;
        }
        const float inv = 1.0f / ((s0 + s1) + (s2 + s3));   // sum >= 1 (the maximum contributes 2^0)
        const uint32_t inv2 = pack_f16x2(inv, inv);
#pragma unroll
        for (int j = 0; j < 16; ++j) { lo[j] = mul_f16x2(lo[j], inv2); hi[j] = mul_f16x2(hi[j], inv2); }
        // P of local head g -> columns [32 g, 32 g + 32) of the set's region: S columns of head g / 2, which this warp
        // consumed (loaded AND waited for) before; the loads in flight address head g + 1 = columns >= 64 (g + 1)
        tmem_st16(t_set + 32 * g, lo);
        tmem_st16(t_set + 32 * g + 16, hi);
        if (more) tmem_ld_wait();
      };
      {
        uint32_t va[32], wa[32], vb[32], wb[32];
        tmem_ld32(t_set, va);
        tmem_ld32(t_set + 32, wa);
        tmem_ld_wait();
        head_probs(va, wa, vb, wb, 0, true);
        head_probs(vb, wb, va, wa, 1, true);
        head_probs(va, wa, vb, wb, 2, true);
        head_probs(vb, wb, va, wa, 3, false);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_mma(&p_full[hs]);
      if (warp == 2 && lane == 0) XA_STAMP(9);

      // ---- epilogue: four output quarters, this warp stores 32-column chunks hs and hs + 2 of each ----
      const int row0 = tile * XA_BM + qd * 32;
      uint8_t* my_stg0 = stg + ew * 2 * 4096;
      int sbuf = 0;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        const int buf = q & 1;
        const uint32_t u = 2u * (uint32_t)it + (uint32_t)(q >> 1);
        mbar_wait(&o_full[buf], u & 1u);
        tc_fence_after();
        if (warp == 2 && lane == 0) XA_STAMP(10 + q);
        const uint32_t t_o = tmem_base + lane_off + (buf ? 384u : 128u);
#pragma unroll 1
        for (int c = hs; c < 4; c += 2) {
          uint32_t o[32];
          tmem_ld32(t_o + 32 * c, o);
          tmem_ld_wait();
          if (c + 2 >= 4) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_mma(&o_empty[buf]);
          }
          const float* bs = s_bias + q * 128 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) + bs[j]);
          uint8_t* my_stg = my_stg0 + sbuf * 4096;
          if (elect_one()) bulk_wait_group_read<1>();   // the store issued two chunks ago has been read (elected lane: owns the bulk groups)
          __syncwarp();
          const uint32_t sdst = smem_u32(my_stg) + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(sdst + ((j ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            // (reading the residual chunk into registers ahead of the accumulator and issuing a plain TMA store was
            // measured slower than the L2-side reduction: 69 vs 56 us per launch)
            tma_reduce_add_2d(&tmO, my_stg, q * 128 + c * 32, row0);
            bulk_commit_group();
          }
          sbuf ^= 1;
        }
      }
      if (warp == 2 && lane == 0) XA_STAMP(14);
    }
    if (elect_one()) bulk_wait_group<0>();
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's smem / barriers / TMEM stay valid until both CTAs are done
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// CTA pairs are OFF by default: parity-green, but measured slower at 64 frames (62.5 vs 56.3 us per launch): the O phase
// is paced by the residual read-modify-write of all CTAs at once and the S phase by L2 -> SM delivery, and the pair
// form shortens neither enough to pay for its longer softmax / epilogue tail. RALD_B200_XATTN_PAIR=1 enables it.
static bool xattn_pair_enabled() {
  const char* e = getenv("RALD_B200_XATTN_PAIR");
  return e != nullptr && e[0] == '1';
}

// h[T][512] += fused cross-attention of xn[T][512] against the folded context operands of one block.
int xattn_fused(const void* xn, const void* kp, const void* vt, const float* bias, float* h, int frames,
                int rows_per_frame, int frame0, int total_frames, cudaStream_t stream) {
  RALD_REQUIRE(xn != nullptr && kp != nullptr && vt != nullptr && h != nullptr, "xattn: null pointer");
  RALD_REQUIRE(frames > 0 && rows_per_frame % XA_BM == 0 && frame0 >= 0 && frame0 + frames <= total_frames,
               "xattn: frames=%d rows/frame=%d frame0=%d total=%d", frames, rows_per_frame, frame0, total_frames);
  const int64_t T = (int64_t)frames * rows_per_frame;
  CUtensorMap tmX, tmK, tmV, tmO;
  RALD_TRY(make_tmap_2d_bf16(&tmX, xn, (uint64_t)T, XA_DIM, XA_DIM, XA_BM));
  RALD_TRY(make_tmap_2d_bf16(&tmK, kp, (uint64_t)8 * total_frames * XA_KEYS, XA_DIM, XA_DIM, XA_KEYS));
  // CTA pairs when the row tiles pair up inside a frame and there are enough of them to fill the machine
  const int sms = device_sm_count();
  const bool pair = xattn_pair_enabled() && (rows_per_frame / XA_BM) % 2 == 0 && T / XA_BM >= sms;
  RALD_TRY(make_tmap_2d_bf16(&tmV, vt, (uint64_t)8 * XA_DIM, (uint64_t)total_frames * XA_KEYS,
                             (uint64_t)total_frames * XA_KEYS, pair ? 64 : 128));
  RALD_TRY(make_tmap_out(&tmO, h, (uint64_t)T, XA_DIM, XA_DIM, true));
  XattnParams p;
  p.bias = bias;
  p.num_tiles = (int)(T / XA_BM);
  p.tiles_per_frame = rows_per_frame / XA_BM;
  p.frame0 = frame0;
  p.total_frames = total_frames;
  p.dbg = g_xattn_dbg;
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(xattn_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM));
    RALD_CHECK_CUDA(cudaFuncSetAttribute(xattn_fused_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, XA_SMEM));
    configured = true;
  }
  // executed flops: S and O products, 2 x (2 * 128 * 512 * 512) per tile
  ProfScope prof(FAM_XATTN, stream, 4.0 * (double)T * XA_DIM * XA_HK);
  if (pair) {
    const int units = p.num_tiles / 2 < sms / 2 ? p.num_tiles / 2 : sms / 2;
    RALD_CHECK_CUDA(launch_pdl_cluster(xattn_fused_kernel<2>, dim3(2 * units), dim3(XA_THREADS), XA_SMEM, stream, 2u, tmX,
                                       tmK, tmV, tmO, p));
  } else {
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    RALD_CHECK_CUDA(launch_pdl(xattn_fused_kernel<1>, dim3(grid), dim3(XA_THREADS), XA_SMEM, stream, tmX, tmK, tmV, tmO, p));
  }
  RALD_LAUNCHED();
  return 0;
}

// The sub-layer as two GEMMs (gemm.cu: grouped K' / VT operands, per-head softmax in the first epilogue). The fused
// kernel works on 128-row tiles with all 8 heads = 4 tiles per frame; below ~32 frames those leave most SMs idle and
// every launch is latency-bound, whereas the GEMM scheduler picks tile widths that cover the machine at any batch.
int xattn_split(const void* xn, const void* kp, const void* vt, const float* bias, float* h, void* probs_f16, int frames,
                int rows_per_frame, int frame0, int total_frames, cudaStream_t stream) {
  RALD_REQUIRE(xn != nullptr && kp != nullptr && vt != nullptr && h != nullptr && probs_f16 != nullptr,
               "xattn_split: null pointer");
  RALD_REQUIRE(frames > 0 && rows_per_frame % (2 * XA_BM) == 0 && frame0 >= 0 && frame0 + frames <= total_frames,
               "xattn_split: frames=%d rows/frame=%d frame0=%d total=%d", frames, rows_per_frame, frame0, total_frames);
  const int T = frames * rows_per_frame;
  RALD_TRY(gemm_xattn_scores(xn, kp, probs_f16, T, XA_DIM, XA_HK, rows_per_frame, frame0, total_frames, stream));
  return gemm_xattn_out(probs_f16, vt, bias, h, T, XA_DIM, XA_HK, rows_per_frame, frame0, total_frames, stream);
}

// Folds attn2.to_q / to_out of every block into the per-frame context operands (once per sample):
//   ctxkv  bf16 [frames*64][depth*1024]   per block K | V projections of the conditioning tokens (both bf16)
//   wq_t   bf16 [depth][512 in][512 q]    attn2.to_q.weight TRANSPOSED and pre-scaled by log2(e)/sqrt(64)
//   w_o    bf16 [depth][512 out][512 in]  attn2.to_out.0.weight
//   kp     bf16 [depth][8][frames][64][512]        vt  fp16 [depth][8][512][frames*64]
// 2 x depth x 8 small-K (K = 64) GEMMs on the tcgen05 kernel.
int xattn_fold(const void* ctxkv, const void* wq_t, const void* w_o, int depth, int frames, void* kp, void* vt,
               cudaStream_t stream) {
  RALD_REQUIRE(ctxkv != nullptr && wq_t != nullptr && w_o != nullptr && kp != nullptr && vt != nullptr,
               "xattn_fold: null pointer");
  RALD_REQUIRE(depth > 0 && frames > 0, "xattn_fold: depth=%d frames=%d", depth, frames);
  const __nv_bfloat16* ctx = reinterpret_cast<const __nv_bfloat16*>(ctxkv);
  const __nv_bfloat16* wq = reinterpret_cast<const __nv_bfloat16*>(wq_t);
  const __nv_bfloat16* wo = reinterpret_cast<const __nv_bfloat16*>(w_o);
  __nv_bfloat16* kpo = reinterpret_cast<__nv_bfloat16*>(kp);
  __nv_bfloat16* vto = reinterpret_cast<__nv_bfloat16*>(vt);
  const int64_t ldc = (int64_t)depth * 2 * XA_DIM;
  const int FK = frames * XA_KEYS;
  for (int n = 0; n < depth; ++n) {
    for (int hd = 0; hd < 8; ++hd) {
      // K'[(f,key)][i] = K[(f,key)][hd*64 + d] . wq_t[i][hd*64 + d]
      RALD_TRY(gemm_bf16(ctx + (int64_t)n * 2 * XA_DIM + hd * 64, ldc, wq + (int64_t)n * XA_DIM * XA_DIM + hd * 64, XA_DIM,
                         kpo + ((int64_t)(n * 8 + hd) * FK) * XA_DIM, XA_DIM, nullptr, nullptr, 0, FK, XA_DIM, 64, 0, 0,
                         stream));
      // VT[o][(f,key)] = w_o[o][hd*64 + d] . V[(f,key)][hd*64 + d]
      // (all columns written as fp16: the O product multiplies fp16 probabilities with fp16 values)
      RALD_TRY(gemm_bf16_f16cols(wo + (int64_t)n * XA_DIM * XA_DIM + hd * 64, XA_DIM,
                                 ctx + (int64_t)n * 2 * XA_DIM + XA_DIM + hd * 64, ldc,
                                 vto + ((int64_t)(n * 8 + hd) * XA_DIM) * FK, FK, nullptr, XA_DIM, FK, 64, 0, 64, stream));
    }
  }
  return 0;
}

}  // namespace rald

// Debug hook like rald_gemm_debug_buffer: CTA 0 stores %globaltimer stamps of its first 4 tiles at dev_buf[tile*16+i]:
// MMA thread 0 tile start, 1 S issued, 2/3 P of head set 0/1 ready, 4..7 O quarter issued; softmax warp 2: 8 S ready,
// 9 P written, 10..13 O quarter ready, 14 tile stored.
extern "C" int rald_xattn_debug_buffer(unsigned long long* dev_buf) {
  rald::g_xattn_dbg = dev_buf;
  return 0;
}

extern "C" int rald_xattn_fold(const void* ctxkv_bf16, const void* wq_t_scaled, const void* w_o, int depth, int frames,
                               void* kp, void* vt, void* stream) {
  return rald::xattn_fold(ctxkv_bf16, wq_t_scaled, w_o, depth, frames, kp, vt, static_cast<cudaStream_t>(stream));
}

extern "C" int rald_xattn_split(const void* xn, const void* kp, const void* vt, const float* bias, float* h,
                                void* probs_f16, int frames, int rows_per_frame, int frame0, int total_frames,
                                void* stream) {
  return rald::xattn_split(xn, kp, vt, bias, h, probs_f16, frames, rows_per_frame, frame0, total_frames,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int rald_xattn_fused(const void* xn, const void* kp, const void* vt, const float* bias, float* h, int frames,
                                int rows_per_frame, int frame0, int total_frames, void* stream) {
  return rald::xattn_fused(xn, kp, vt, bias, h, frames, rows_per_frame, frame0, total_frames,
                           static_cast<cudaStream_t>(stream));
}
