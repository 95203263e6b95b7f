// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//     out[M, N] = epilogue( A[M, K] (bf16, K-major) x W[N, K]^T (bf16, K-major) ), fp32 accumulate in TMEM.
//
//   warp 0 (1 lane)  TMA producer: 128 x 64 A tile + BN x 64 W tile per stage, 128B-swizzled smem ring
//   warp 1 (1 lane)  tcgen05.mma issuer (M=128, N=BN, K=16 per instruction), accumulators double-buffered in TMEM
//   warps 2..5       epilogue: tcgen05.ld -> registers -> (+bias) / GEGLU -> 128B-swizzled smem staging ->
//                    TMA tensor store (bf16 / fp32) or TMA reduce-add (fp32 `out += tile`: the residual add
//                    h += f(h) happens in the L2, the residual is never read by the SM)
//
// Epilogue latency matters as much as the main loop here (K = 512 for most layers: 8 k-blocks per tile): the bias
// of a tile is fetched while its MMAs run, all global traffic of the epilogue is asynchronous (TMA), and the TMEM
// accumulator is released as soon as its last column has been read so the next tile's MMAs overlap the stores.
// A generic direct-store epilogue remains for the rarely used combinations (residual that is not in place or a
// broadcast table, bf16 output narrower than a 128-byte row).
//
// Replaces, for every nn.Linear on the reference hot path, the cuBLAS sgemm + separate elementwise ops:
//   model/models_radar_generation.py:58-64 (to_q/to_k/to_v), :76 (to_out + residual :166-168),
//   :91-95 (GEGLU proj + gelu gate), :113 (ff out); model/models_ae.py:60-62, 78-80, 87-105.
#include <stdlib.h>

#include "host.cuh"
#include "ptx.cuh"
#include "kernels.h"

namespace rald {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps
constexpr int GEMM_EPI_WARPS = 8;
constexpr int EPI_GENERIC = 0, EPI_TMA_STORE = 1, EPI_TMA_REDUCE = 2;
constexpr int STG_BYTES = 32 * 128;  // one staging tile: 32 rows x 128 bytes

// CG = 1: one CTA computes a 128 x BN tile. CG = 2: a CTA pair (cta_group::2, the two SMs of a TPC) computes a
// 256 x BN tile with ONE 256-row MMA per K step; each CTA stages its 128 rows of A and only HALF of the W tile, so
// the shared-memory fill traffic per flop drops by a third and the ring gets deeper.
// Variants measured and rejected in round 1 (A-resident K = 512 form, 64-row tiles, 128-row TMA stores, a second TMA
// producer warp, L2 prefetch of the next GEMM's weights, 16 ring stages) are documented in DESIGN.md §7 and live in
// the git history (commit 8837a6c); they are no longer compiled.
template <int BN, int CG = 1>
struct GemmCfg {
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = (BN / CG) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_TOTAL = GEMM_EPI_WARPS * STG_BYTES;   // one staging tile per epilogue warp
  static constexpr int BIAS_BYTES = 2 * BN * 4;
  // no alignment slack: dynamic shared memory starts 1024-byte aligned (checked at kernel entry)
  static constexpr int FIXED = 512 /*barriers*/ + STG_TOTAL + BIAS_BYTES;
  static constexpr int STAGES_RAW = (232448 - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static_assert(2 * STAGES + 7 <= 64, "barrier block too small");
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator buffers
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + FIXED;
};

struct GemmParams {
  void* out;
  int64_t ldo;
  const float* bias;   // [N] or null (GEGLU: packed order, see geglu_pack_index in the Python runtime)
  const float* resid;  // [M, ldr] fp32 or null (generic epilogue only)
  int64_t ldr;
  int64_t resid_mod;   // > 0: residual row = row % resid_mod (a [resid_mod, N] table broadcast over frames)
  int M, N, K;
  int a_k;             // columns of A. a_k == K: plain GEMM. a_k == K / 2: SPLIT WEIGHTS — W holds [W_hi | W_lo]
                       // (two bf16 matrices whose sum carries 16 mantissa bits) along K and the A tile of k-block kb
                       // is read at column (kb * 64) % a_k, i.e. out = A W_hi^T + A W_lo^T in one accumulator
  int num_m_blks, num_n_blks;
  int f16_start, f16_period;  // bf16 mode: output columns with (col % f16_period) >= f16_start are written as fp16
  int w_static;               // W does not depend on the preceding kernel: its first ring stages are loaded BEFORE the
                              // programmatic-dependency wait (hides the DRAM latency of the weights behind the
                              // predecessor's tail in the latency-bound small-batch regime); CG = 1 only
  int gelu_exact;             // GEGLU epilogue: erf GELU (A&S 7.1.26, fp32 round-off level) instead of the logistic fit
  // Per-frame B operands of the folded cross-attention (xattn.cu: K' / VT built once per sample by xattn_fold). Rows
  // [m, m + grp_rows) of A belong to frame grp_frame0 + m / grp_rows of grp_total frames:
  //   b_mode 1  scores S = xn K'^T: W = K' of ONE block, bf16 [8 heads][grp_total][64 keys][K]; the B tile of column
  //             block n_blk is one 64-row box per head (heads n_blk * BN / 64 ...), rows ((head * grp_total + f) * 64
  //   b_mode 2  output O = P VT^T: W = VT of one block, fp16 [8 heads][N][grp_total * 64]; k-block kb is head kb, its B
  //             tile rows kb * N + n_blk * BN at columns f * 64
  int b_mode, grp_rows, grp_frame0, grp_total;
  int ab_f16;                 // A and W are IEEE fp16 (kind::f16 with fp16 operands) instead of bf16
  int w_k_off;                // added to the K coordinate of every W tile (A unaffected): W is read `w_k_off` columns to
                              // the right of A (negative: left; columns outside [0, K) read zeros) — a tap of the
                              // convolution weight gradient on the flattened padded voxel grid (enc_bwd.cu)
  int tap_n;                  // > 0: the N extent is n_taps blocks of tap_n columns that all read the SAME tap_n rows of W, block
  int tap_shift[16];          // t with column shift tap_shift[t] instead of w_k_off (all taps of a convolution weight gradient
                              // in ONE launch). tap_n < BN: a W tile is BN / tap_n boxes of tap_n rows, one per tap, each with
                              // its own shift — several taps share one A tile and one wide MMA
  int panel_len, b_halo;      // panel_len > 0: A is [K / panel_len panels][M][panel_len] and W [panels][w rows][panel_len + 2 b_halo]
                              // (K-panel-major: the rows of a tile lie within a few pages; with a 5.5 MB row pitch every row of
                              // a box is in another 2 MB page and the same GEMM runs 3x slower, tools/gpu_pitch_probe.py). W
                              // panels carry b_halo columns of their neighbours on both sides so that shifted boxes stay inside
  int k_splits;               // > 1 (fp32 reduce-add epilogue, no bias): every output tile is computed by k_splits work items,
  int kb_per_split;           // each over kb_per_split k-blocks, all adding into `out` — the weight-gradient GEMMs of the
                              // training step (few output tiles, K = rows of the batch). Summation order across the splits
                              // is not fixed (fp32 adds in L2): results reproduce to ~1e-7 relative, not bit for bit
  unsigned long long* dbg;  // optional [gridDim.x][8] %globaltimer stamps of the first tile (tools/gemm_phases.py)
};

static unsigned long long* g_gemm_dbg = nullptr;
static thread_local int g_w_static = 0;               // GemmStaticWeights scopes alive on this thread
#define GEMM_STAMP(slot)                                                        \
  do {                                                                          \
    if (p.dbg != nullptr) p.dbg[blockIdx.x * 8 + (slot)] = global_timer_ns();   \
  } while (0)

// ---- generic epilogue for one 32-column chunk already in registers (direct global loads / stores) ----
template <int OUT_MODE>
__device__ __forceinline__ void epilogue_direct(uint32_t (&v)[32], const GemmParams& p, int64_t row, bool row_ok,
                                                int col0) {
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
      v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
      v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
      v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
    }
  }
  if (OUT_MODE != 2 && p.resid != nullptr && row_ok) {
    const int64_t rrow = p.resid_mod > 0 ? row % p.resid_mod : row;
    const float4* r4 = reinterpret_cast<const float4*>(p.resid + rrow * p.ldr + col0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 r = r4[j];
      v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + r.x);
      v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + r.y);
      v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + r.z);
      v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + r.w);
    }
  }
  if (OUT_MODE == 1) {
    if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  } else if (OUT_MODE == 0) {
    if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + col0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dst[j] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])),
                            pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])),
                            pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])),
                            pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])));
      }
    }
  } else {
    // GEGLU: within every packed 32-column group, columns [0,16) are the value half and [16,32) the gate half
    // of the same 16 output features (reference: x, gate = proj(x).chunk(2); x * gelu(gate)).
    float o[16];
    if (p.gelu_exact) {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) * gelu_erf(__uint_as_float(v[16 + j]));
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) * gelu_act(__uint_as_float(v[16 + j]));
    }
    if (row_ok) {
      uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + (col0 >> 1));
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        dst[j] = make_uint4(pack_bf16x2(o[8 * j + 0], o[8 * j + 1]), pack_bf16x2(o[8 * j + 2], o[8 * j + 3]),
                            pack_bf16x2(o[8 * j + 4], o[8 * j + 5]), pack_bf16x2(o[8 * j + 6], o[8 * j + 7]));
      }
    }
  }
}

// OUT_MODE: 0 = bf16 [M,N]; 1 = fp32 [M,N]; 2 = GEGLU -> bf16 [M,N/2]; 3 = softmax over every group of 64 columns
// (scores already in log2 units) -> fp16 probabilities [M,N] (TMA-store epilogue only)
// SPLITK: the weight-gradient form (K splits, W column shifts / taps). A separate instantiation so that the forward
// kernels keep their exact instruction stream (no runtime division by k_splits, no shift logic in the producer).
template <int BN, int OUT_MODE, int EPI, int CG, int G, bool SPLITK = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  using Cfg = GemmCfg<BN, CG>;
  // G k-blocks share one ring slot = ONE full / empty barrier pair: the MMA thread then probes an mbarrier once per
  // 4 G MMAs instead of once per 4. Measured (tools/micro/mma_rate.cu): a try_wait on an ALREADY COMPLETE mbarrier
  // between groups of 4 MMAs costs the issuing thread ~260 cycles (490 cycles per 64-wide k-block in all, whatever
  // N) — more than the 180 cycles of tensor work of a k-block at N = 32 and barely less than the 512 at N = 256.
  static_assert(Cfg::STAGES % G == 0 || G == 1, "ring depth must be a multiple of the group");
  constexpr int SLOTS = Cfg::STAGES / G;
  constexpr int SLOT_BYTES = G * Cfg::STAGE_BYTES;
  constexpr int BM = GEMM_BM;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;   // 0 = leader of the pair (issues the MMAs)
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // tile-processing unit (CTA or CTA pair)
  const int num_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int TILE_M = BM * CG;
  constexpr int STAGES = Cfg::STAGES;
  // accumulator columns consumed per staged 128-byte output row
  constexpr int CHUNK = OUT_MODE == 1 ? 32 : (OUT_MODE == 2 ? 128 : 64);
  static_assert(OUT_MODE != 3 || EPI == EPI_TMA_STORE, "softmax epilogue: TMA store only");
  static_assert(EPI == EPI_GENERIC || BN % CHUNK == 0, "tile narrower than one staged output row");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // 128-byte swizzle atoms need 1024-byte aligned tiles
  uint8_t* stg = smem + STAGES * Cfg::STAGE_BYTES;                    // 1024-byte aligned (stage sizes are)
  float* s_bias = reinterpret_cast<float*>(stg + Cfg::STG_TOTAL);     // [2][BN]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_bias) + Cfg::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.num_m_blks * p.num_n_blks * (SPLITK ? p.k_splits : 1);
  const int num_kb_all = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int num_kb = (SPLITK && p.k_splits > 1) ? p.kb_per_split : num_kb_all;   // k-blocks per work item
  // work item -> (row block, column block, first k-block); without K splits a work item is an output tile
  auto decode = [&](int tile, int& m_blk, int& n_blk, int& kb0) {
    int t2 = tile;
    kb0 = 0;
    if (SPLITK) {
      t2 = tile / p.k_splits;
      kb0 = (tile - t2 * p.k_splits) * num_kb;
    }
    m_blk = t2 / p.num_n_blks;
    n_blk = t2 - m_blk * p.num_n_blks;
  };
  // tile walk of this unit: tiles unit, unit + num_units, ...
  const int t_begin = unit, t_end = num_tiles, t_step = num_units;
  // A column of k-block kb (split weights: the A tiles are re-read for the second half of K)
  auto a_col = [&](int kb) { const int c = kb * GEMM_BK; return c >= p.a_k ? c - p.a_k : c; };
  // the A tile of k-block kb, rows row0 ...
  auto load_a = [&](uint8_t* sa, uint64_t* bar, int kb, int row0) {
    if (SPLITK && p.panel_len > 0) {
      const int c = kb * GEMM_BK, pn = c / p.panel_len;
      tma_load_3d(sa, &tmA, bar, c - pn * p.panel_len, row0, pn);
    } else {
      tma_load_2d(sa, &tmA, bar, a_col(kb), row0);
    }
  };
  // this CTA's share of the W tile of (k-block kb, tile (m_blk, n_blk)) -> sb, credited to `bar`
  auto load_b = [&](uint8_t* sb, uint64_t* bar, int kb, int m_blk, int n_blk) {
    auto ld = [&](uint8_t* dst, int x, int y) {
      if (CG == 2) tma_load_2d_pair(dst, &tmB, bar, x, y);
      else tma_load_2d(dst, &tmB, bar, x, y);
    };
    if (p.b_mode == 0) {
      int row = n_blk * BN + (int)cta_rank * (BN / CG), shift = 0;
      if (SPLITK) {
        shift = p.w_k_off;
        if (p.panel_len > 0) {
          // K-panel-major operands (3-D tensor maps: column in panel, row, panel)
          const int c = kb * GEMM_BK, pn = c / p.panel_len, cc = c - pn * p.panel_len + p.b_halo;
          if (p.tap_n > 0 && p.tap_n < BN) {
            const int tap0 = row / p.tap_n;
            for (int i = 0; i < BN / p.tap_n; ++i)
              tma_load_3d(sb + (size_t)i * p.tap_n * 128, &tmB, bar, cc + p.tap_shift[tap0 + i], 0, pn);
          } else if (p.tap_n > 0) {
            const int tap = row / p.tap_n;
            tma_load_3d(sb, &tmB, bar, cc + p.tap_shift[tap], row - tap * p.tap_n, pn);
          } else {
            tma_load_3d(sb, &tmB, bar, cc + shift, row, pn);
          }
          return;
        }
        if (p.tap_n > 0 && p.tap_n < BN) {
          // several taps per tile: one box of tap_n rows (all rows of W) per tap, stacked in the B tile
          const int tap0 = row / p.tap_n;
          for (int i = 0; i < BN / p.tap_n; ++i)
            ld(sb + (size_t)i * p.tap_n * 128, kb * GEMM_BK + p.tap_shift[tap0 + i], 0);
          return;
        }
        if (p.tap_n > 0) {
          const int tap = row / p.tap_n;
          shift = p.tap_shift[tap];
          row -= tap * p.tap_n;
        }
      }
      ld(sb, kb * GEMM_BK + shift, row);
    } else {
      const int f = p.grp_frame0 + (m_blk * TILE_M) / p.grp_rows;
      if (p.b_mode == 1) {
        constexpr int HB = (BN / CG) / 64 > 0 ? (BN / CG) / 64 : 1;   // heads in this CTA's share of the tile
#pragma unroll
        for (int i = 0; i < HB; ++i) {
          const int head = n_blk * (BN / 64) + (int)cta_rank * HB + i;
          ld(sb + i * (64 * 128), kb * GEMM_BK, (head * p.grp_total + f) * 64);
        }
      } else {
        ld(sb, f * 64, kb * p.N + n_blk * BN + (int)cta_rank * (BN / CG));
      }
    }
  };

  if (threadIdx.x == 0) {
    GEMM_STAMP(0);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI != EPI_GENERIC) tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], GEMM_EPI_WARPS * CG);  // one arrive per epilogue warp (of both CTAs)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish2();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int w_pre = 0;            // ring stages whose W tile is already in flight (producer warp only; warp-uniform)
  if (CG == 1 && p.w_static != 0 && warp == 0) {
    // static weights: the W tiles of this CTA's first SLOTS ring slots do not depend on the preceding kernel
    int tile = t_begin, kb = 0;
    while (w_pre < SLOTS && tile < t_end) {
      int m_blk, n_blk, kb0;
      decode(tile, m_blk, n_blk, kb0);   // (k_splits == 1 here: w_static is off for split-K launches)
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[w_pre], SLOT_BYTES);
#pragma unroll
        for (int g = 0; g < G; ++g)
          load_b(smem + w_pre * SLOT_BYTES + g * Cfg::STAGE_BYTES + Cfg::A_BYTES, &full_bar[w_pre], kb + g, m_blk, n_blk);
      }
      __syncwarp();
      ++w_pre;
      kb += G;
      if (kb == num_kb) { kb = 0; tile += t_step; }
    }
  }
  // A (and an in-place residual) come from the preceding kernel
  pdl_wait();
  pdl_launch_dependents();  // the next kernel may set itself up on idle SMs while this one runs

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (whole warp, one elected lane issues: see the MMA issuer — a TMA load under `lane == 0` costs the same
    // ELECT / R2UR.BROADCAST loop per instruction)
    {
      if (elect_one()) GEMM_STAMP(1);
      int s = 0;
      uint32_t ph = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step) {
        int m_blk, n_blk, kb0;
        decode(tile, m_blk, n_blk, kb0);
        for (int kb = kb0; kb < kb0 + num_kb; kb += G) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* slot = smem + s * SLOT_BYTES;
          if (elect_one()) {
            if (CG == 2) {
              // each CTA loads its own 128 rows of A and its half of the W tile; every byte of the pair is credited
              // to the LEADER's full barrier, on which only the leader arrives
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * SLOT_BYTES);
#pragma unroll
              for (int g = 0; g < G; ++g) {
                uint8_t* sa = slot + g * Cfg::STAGE_BYTES;
                tma_load_2d_pair(sa, &tmA, &full_bar[s], a_col(kb + g), m_blk * TILE_M + (int)cta_rank * BM);
                load_b(sa + Cfg::A_BYTES, &full_bar[s], kb + g, m_blk, n_blk);
              }
            } else if (w_pre > 0) {
              // transaction bytes announced and W tiles issued before the dependency wait: only A is left
#pragma unroll
              for (int g = 0; g < G; ++g)
                load_a(slot + g * Cfg::STAGE_BYTES, &full_bar[s], kb + g, m_blk * BM);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], SLOT_BYTES);
#pragma unroll
              for (int g = 0; g < G; ++g) {
                uint8_t* sa = slot + g * Cfg::STAGE_BYTES;
                load_a(sa, &full_bar[s], kb + g, m_blk * BM);
                load_b(sa + Cfg::A_BYTES, &full_bar[s], kb + g, m_blk, n_blk);
              }
            }
          }
          __syncwarp();
          if (CG != 2 && w_pre > 0) --w_pre;
          if (++s == SLOTS) {
            s = 0;
            if (ph == 0 && elect_one()) GEMM_STAMP(7);   // first pass over the ring issued (tools/gemm_phases.py: "ring-issued")
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp runs this loop with warp-uniform values and one ELECTED lane issues. Under `lane == 0` the compiler
    // cannot keep descriptors and addresses in uniform registers: it wrapped every tcgen05.mma / commit in an ELECT /
    // R2UR.BROADCAST loop of ~14 instructions (cuobjdump: 3.8 R2UR.BROADCAST per UTCHMMA), ~120 clocks per MMA — hidden
    // behind a 256-wide MMA, but the pace of every narrower tile (the "0.25 us per k-block whatever the shape" of the
    // small-batch regime and the "260-clock mbarrier probe" of round 1's microbenchmark were this loop).
    if (cta_rank == 0) {
      const uint32_t idesc = p.ab_f16 ? make_idesc(FMT_F16, TILE_M, BN, 0, 0) : make_idesc(FMT_BF16, TILE_M, BN, 0, 0);
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = t_begin; tile < t_end; tile += t_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base_u + acc * BN;
        for (int kb = 0; kb < num_kb; kb += G) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (elect_one()) {
            if (it == 0 && kb == 0) GEMM_STAMP(2);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const uint32_t sa = smem_u + s * SLOT_BYTES + g * Cfg::STAGE_BYTES;
              const uint64_t a_desc = make_sdesc_sw128(sa, 16, 1024);
              const uint64_t b_desc = make_sdesc_sw128(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
              for (int k = 0; k < GEMM_BK / 16; ++k) {
                // +32 bytes (= 2 in the >>4 address field) per K=16 step inside the 128-byte swizzle row
                const uint32_t accum = ((kb + g) | k) != 0 ? 1u : 0u;
                if (CG == 2) mma_f16_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accum);
                else mma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, accum);
              }
            }
            // frees this ring slot (in both CTAs of a pair) once the MMAs above have read it
            if (CG == 2) tc_commit_pair(&empty_bar[s]);
            else tc_commit(&empty_bar[s]);
          }
          __syncwarp();
          if (++s == SLOTS) { s = 0; ph ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (elect_one()) {
          if (CG == 2) tc_commit_pair(&tmem_full_bar[acc]);
          else tc_commit(&tmem_full_bar[acc]);
          if (it == 0) GEMM_STAMP(3);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quarter: warp pair member `hs` takes the even / odd column chunks of the tile.
    const int q = warp & 3;          // TMEM lane quarter this warp is allowed to touch
    const int ew = warp - 2;         // 0..7
    const int hs = ew >> 2;          // 0: even chunks, 1: odd chunks
    const int row_in_tile = q * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..255 among the epilogue threads
    int it = 0;
    for (int tile = t_begin; tile < t_end; tile += t_step, ++it) {
      int m_blk, n_blk, kb0_unused;
      decode(tile, m_blk, n_blk, kb0_unused);
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      float* bias_s = s_bias + acc * BN;
      // the accumulator is handed back on the LEADER's barrier (its MMA thread waits for both CTAs' epilogues)
      auto release_acc = [&]() {
        if (CG == 2) mbar_arrive_leader(&tmem_empty_bar[acc]);
        else mbar_arrive(&tmem_empty_bar[acc]);
      };
      if (EPI != EPI_GENERIC) {
        // bias of this tile -> smem while the MMAs of the tile are still running
        for (int j = et; j < BN; j += 32 * GEMM_EPI_WARPS) {
          const int col = n_blk * BN + j;
          bias_s[j] = (p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      if (it == 0 && threadIdx.x == 64) GEMM_STAMP(4);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      if (EPI == EPI_GENERIC) {
        const int64_t row = static_cast<int64_t>(m_blk) * TILE_M + cta_rank * BM + row_in_tile;
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int c = hs; c < BN / 32; c += 2) {
          const int col0 = n_blk * BN + c * 32;
          if (col0 >= p.N) break;
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          epilogue_direct<OUT_MODE>(v, p, row, row_ok, col0);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) release_acc();
      } else {
        constexpr int NCHUNK = BN / CHUNK;
        const int row0 = m_blk * TILE_M + (int)cta_rank * BM + q * 32;
        if (NCHUNK == 1 && hs == 1) {
          // a single 128-byte output row per tile: the odd warps have no chunk, they only release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) release_acc();
        }
#pragma unroll 1
        for (int c = hs; c < NCHUNK; c += 2) {
          const int acol0 = c * CHUNK;                 // accumulator column inside the tile
          const bool live = n_blk * BN + acol0 < p.N;  // chunks past N are read (uniform TMEM hand-off) but not stored
          uint32_t o[32];                              // one 128-byte output row
          if (OUT_MODE == 1) {
            tmem_ld32(t_row + acol0, o);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) + bias_s[acol0 + j]);
          } else if (OUT_MODE == 0) {
            // 64-column chunks never straddle an fp16 / bf16 boundary (boundaries are multiples of 64 columns)
            const bool as_f16 = p.f16_period > 0 && ((n_blk * BN + acol0) % p.f16_period) >= p.f16_start;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint32_t v[32];
              tmem_ld32(t_row + acol0 + 32 * h, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float lo = __uint_as_float(v[2 * j]) + bias_s[acol0 + 32 * h + 2 * j];
                const float hi = __uint_as_float(v[2 * j + 1]) + bias_s[acol0 + 32 * h + 2 * j + 1];
                o[16 * h + j] = as_f16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
              }
            }
          } else if (OUT_MODE == 3) {
            // per-head softmax of the folded cross-attention: the 64 columns of this chunk are one head's scores, already
            // in log2 units; the arithmetic (and its order) is that of xattn_fused_kernel's head_probs, so both forms of
            // the sub-layer produce the same probabilities bit for bit
            uint32_t v[32], w[32];
            tmem_ld32(t_row + acol0, v);
            tmem_ld32(t_row + acol0 + 32, w);
            tmem_ld_wait();
            float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]), m2 = __uint_as_float(w[0]),
                  m3 = __uint_as_float(w[1]);
#pragma unroll
            for (int j = 2; j < 32; j += 2) {
              m0 = fmaxf(m0, __uint_as_float(v[j])); m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
              m2 = fmaxf(m2, __uint_as_float(w[j])); m3 = fmaxf(m3, __uint_as_float(w[j + 1]));
            }
            const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              o[j] = pack_f16x2(ex2_f32(__uint_as_float(v[2 * j]) - mx), ex2_f32(__uint_as_float(v[2 * j + 1]) - mx));
              o[16 + j] = pack_f16x2(ex2_f32(__uint_as_float(w[2 * j]) - mx), ex2_f32(__uint_as_float(w[2 * j + 1]) - mx));
              const float2 a = unpack_f16x2(o[j]), b = unpack_f16x2(o[16 + j]);
              s0 += a.x; s1 += a.y; s2 += b.x; s3 += b.y;
            }
            const float inv = 1.0f / ((s0 + s1) + (s2 + s3));   // sum >= 1 (the maximum contributes 2^0)
            const uint32_t inv2 = pack_f16x2(inv, inv);
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = mul_f16x2(o[j], inv2);
          } else {
            const bool exact = p.gelu_exact != 0;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              uint32_t v[32];
              tmem_ld32(t_row + acol0 + 32 * h, v);
              tmem_ld_wait();
              const float* bs = bias_s + acol0 + 32 * h;
              if (exact) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float a0 = (__uint_as_float(v[2 * j]) + bs[2 * j]) *
                                   gelu_erf(__uint_as_float(v[16 + 2 * j]) + bs[16 + 2 * j]);
                  const float a1 = (__uint_as_float(v[2 * j + 1]) + bs[2 * j + 1]) *
                                   gelu_erf(__uint_as_float(v[16 + 2 * j + 1]) + bs[16 + 2 * j + 1]);
                  o[8 * h + j] = pack_bf16x2(a0, a1);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float a0 = (__uint_as_float(v[2 * j]) + bs[2 * j]) *
                                   gelu_act(__uint_as_float(v[16 + 2 * j]) + bs[16 + 2 * j]);
                  const float a1 = (__uint_as_float(v[2 * j + 1]) + bs[2 * j + 1]) *
                                   gelu_act(__uint_as_float(v[16 + 2 * j + 1]) + bs[16 + 2 * j + 1]);
                  o[8 * h + j] = pack_bf16x2(a0, a1);
                }
              }
            }
          }
          if (c + 2 >= NCHUNK) {
            // every TMEM read of this accumulator by this warp has completed -> hand it back to the MMA warp now
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc();
          }
          if (live) {
            // (the elected lane — always the same one for the full mask — owns this warp's bulk groups; `lane == 0` would
            // cost an ELECT / R2UR.BROADCAST loop per TMA store)
            if (elect_one()) bulk_wait_group_read<0>();  // the staging tile about to be overwritten was read
            __syncwarp();
            const uint32_t sdst = smem_u32(stg + ew * STG_BYTES) + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(sdst + ((j ^ (lane & 7)) << 4), o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              const int ocol = OUT_MODE == 2 ? ((n_blk * BN + acol0) >> 1) : (n_blk * BN + acol0);
              const void* src = stg + ew * STG_BYTES;
              if (EPI == EPI_TMA_REDUCE) tma_reduce_add_2d(&tmO, src, ocol, row0);
              else tma_store_2d(&tmO, src, ocol, row0);
              bulk_commit_group();
            }
          }
        }
      }
      if (it == 0 && threadIdx.x == 64) GEMM_STAMP(5);
    }
    if (EPI != EPI_GENERIC && elect_one()) bulk_wait_group<0>();  // smem must outlive the last store
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's smem / barriers / TMEM stay valid until both CTAs are done
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    if (lane == 0) GEMM_STAMP(6);
  }
}

template <int BN, int OUT_MODE, int EPI, int CG, int G, bool SPLITK = false>
static int launch_gemm_g(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmParams& p,
                         int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG>;
  static_assert(Cfg::STAGES >= 3, "ring too shallow");
  auto kern = gemm_bf16_kernel<BN, OUT_MODE, EPI, CG, G, SPLITK>;
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int num_tiles = p.num_m_blks * p.num_n_blks * (SPLITK ? p.k_splits : 1);   // work items: tiles (x K splits)
  const int max_units = max_ctas / CG;
  const int units = num_tiles < max_units ? num_tiles : max_units;
  ProfScope prof(FAM_GEMM, stream, 2.0 * p.M * p.N * p.K);
  RALD_CHECK_CUDA(launch_pdl_cluster(kern, dim3(units * CG), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, (unsigned)CG,
                                     tmA, tmB, tmO, p));
  RALD_LAUNCHED();
  return 0;
}

// environment switches of this file, read once per process
struct GemmEnv {
  bool group, pair, wpre;
  GemmEnv() {
    auto on = [](const char* name) { const char* e = getenv(name); return e == nullptr || e[0] != '0'; };
    group = on("RALD_B200_GEMM_GROUP");   // =0: one k-block per ring slot
    pair = on("RALD_B200_GEMM_PAIR");     // =0: no CTA-pair tiles
    wpre = on("RALD_B200_GEMM_WPRE");     // =0: no early weight loads
  }
};
static const GemmEnv& gemm_env() {
  static const GemmEnv env;
  return env;
}

// Picks the ring-slot group (k-blocks per full / empty barrier pair, see the kernel): the largest the ring depth
// allows when K is a multiple of it.
template <int BN, int OUT_MODE, int EPI, int CG = 1>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const GemmParams& p,
                       int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr int GMAX = EPI == EPI_GENERIC ? 1
                       : (Cfg::STAGES >= 8 && Cfg::STAGES % 4 == 0) ? 4
                       : (Cfg::STAGES >= 4 && Cfg::STAGES % 2 == 0) ? 2 : 1;
  const int num_kb = p.k_splits > 1 ? p.kb_per_split : (p.K + GEMM_BK - 1) / GEMM_BK;
  if constexpr (OUT_MODE == 1 && EPI == EPI_TMA_REDUCE && CG == 1) {
    // the weight-gradient form is only reachable through the fp32 accumulate-into-out GEMMs
    if (p.k_splits > 1 || p.tap_n > 0 || p.w_k_off != 0 || p.panel_len > 0) {
      // one k-block per ring slot: these GEMMs stream hundreds of k-blocks per work item, and a slot that is refilled
      // as soon as ONE k-block has been consumed keeps more loads in flight than grouped slots (measured, tools/
      // gpu_time_wgrad.py: dW of FF2 70 -> 62 us, of the 512 x 512 projections 22.6 -> 19.2 us at K = 32 768)
      return launch_gemm_g<BN, OUT_MODE, EPI, CG, 1, true>(tmA, tmB, tmO, p, max_ctas, stream);
    }
  }
  if (GMAX > 1 && gemm_env().group && num_kb % GMAX == 0)
    return launch_gemm_g<BN, OUT_MODE, EPI, CG, GMAX>(tmA, tmB, tmO, p, max_ctas, stream);
  return launch_gemm_g<BN, OUT_MODE, EPI, CG, 1>(tmA, tmB, tmO, p, max_ctas, stream);
}

GemmStaticWeights::GemmStaticWeights() { ++g_w_static; }
GemmStaticWeights::~GemmStaticWeights() { --g_w_static; }

struct GemmOpts {
  int64_t resid_mod = 0;
  int f16_start = 0, f16_period = 0;
  int w_split = 0;      // W is [N][2 K] = [W_hi | W_lo]
  int gelu_exact = 0;
  int b_mode = 0, grp_rows = 0, grp_frame0 = 0, grp_total = 0;   // folded cross-attention operands (GemmParams)
  int ab_f16 = 0;
  int split_k = 0;      // allow K splits (fp32 accumulate-into-out GEMMs without bias: weight gradients)
  int w_k_off = 0;      // column shift of the W operand (GemmParams::w_k_off)
  int panel_len = 0, b_halo = 0, n_panels = 0;   // K-panel-major operands (GemmParams::panel_len)
  int tap_n = 0, n_taps = 0;   // N = n_taps * tap_n, W has tap_n rows, tap t is shifted by tap_shift[t]
  int tap_shift[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

static int gemm_impl(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
                     const float* resid, int64_t ldr, int M, int N, int K, int out_mode, int bn_hint,
                     const GemmOpts& o, cudaStream_t stream);

int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
              const float* resid, int64_t ldr, int M, int N, int K, int out_mode, int bn_hint,
              cudaStream_t stream) {
  return gemm_impl(A, lda, W, ldw, out, ldo, bias, resid, ldr, M, N, K, out_mode, bn_hint, GemmOpts(), stream);
}

int gemm_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
                 const float* resid, int64_t ldr, int64_t resid_mod, int M, int N, int K, int out_mode, int bn_hint,
                 cudaStream_t stream) {
  GemmOpts o;
  o.resid_mod = resid_mod;
  return gemm_impl(A, lda, W, ldw, out, ldo, bias, resid, ldr, M, N, K, out_mode, bn_hint, o, stream);
}

int gemm_bf16_accum_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, int w_k_off, float* out, int64_t ldo,
                           int M, int N, int K, cudaStream_t stream) {
  GemmOpts o;
  o.split_k = 1;
  o.w_k_off = w_k_off;
  return gemm_impl(A, lda, W, ldw, out, ldo, nullptr, out, ldo, M, N, K, 1, 0, o, stream);
}

int gemm_bf16_accum_taps(const void* A, int64_t lda, const void* W, int64_t ldw, int w_rows, int n_taps,
                         const int* tap_shifts, int panel_len, int w_halo, float* out, int64_t ldo, int M, int K,
                         cudaStream_t stream) {
  RALD_REQUIRE(n_taps >= 1 && n_taps <= 16 && tap_shifts != nullptr, "gemm taps: 1..16 taps");
  GemmOpts o;
  o.split_k = 1;
  o.tap_n = w_rows;
  o.n_taps = n_taps;
  if (panel_len > 0) {
    RALD_REQUIRE(K % panel_len == 0, "gemm taps: K=%d must be a multiple of the panel length %d", K, panel_len);
    o.panel_len = panel_len;
    o.b_halo = w_halo;
    o.n_panels = K / panel_len;
    for (int t = 0; t < n_taps; ++t)
      RALD_REQUIRE(tap_shifts[t] >= -w_halo && tap_shifts[t] <= w_halo, "gemm taps: shift %d exceeds the halo %d",
                   tap_shifts[t], w_halo);
  }
  for (int t = 0; t < n_taps; ++t) o.tap_shift[t] = tap_shifts[t];
  return gemm_impl(A, lda, W, ldw, out, ldo, nullptr, out, ldo, M, n_taps * w_rows, K, 1, 0, o, stream);
}

int gemm_bf16_f16cols(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
                      int M, int N, int K, int f16_start, int f16_period, cudaStream_t stream) {
  GemmOpts o;
  o.f16_start = f16_start;
  o.f16_period = f16_period;
  return gemm_impl(A, lda, W, ldw, out, ldo, bias, nullptr, 0, M, N, K, 0, 0, o, stream);
}

int gemm_bf16_wsplit(const void* A, int64_t lda, const void* W_hilo, int64_t ldw, void* out, int64_t ldo,
                     const float* bias, const float* resid, int64_t ldr, int M, int N, int K, int out_mode,
                     int f16_start, int f16_period, int gelu_exact, cudaStream_t stream) {
  GemmOpts o;
  o.f16_start = f16_start;
  o.f16_period = f16_period;
  o.w_split = 1;
  o.gelu_exact = gelu_exact;
  return gemm_impl(A, lda, W_hilo, ldw, out, ldo, bias, resid, ldr, M, N, K, out_mode, 0, o, stream);
}

int gemm_xattn_scores(const void* xn, const void* kp_block, void* probs_f16, int T, int dim, int heads_keys,
                      int rows_per_frame, int frame0, int total_frames, cudaStream_t stream) {
  GemmOpts o;
  o.b_mode = 1;
  o.grp_rows = rows_per_frame;
  o.grp_frame0 = frame0;
  o.grp_total = total_frames;
  return gemm_impl(xn, dim, kp_block, dim, probs_f16, heads_keys, nullptr, nullptr, 0, T, heads_keys, dim, 3, 0, o, stream);
}

int gemm_xattn_out(const void* probs_f16, const void* vt_block, const float* bias, float* h, int T, int dim,
                   int heads_keys, int rows_per_frame, int frame0, int total_frames, cudaStream_t stream) {
  GemmOpts o;
  o.b_mode = 2;
  o.grp_rows = rows_per_frame;
  o.grp_frame0 = frame0;
  o.grp_total = total_frames;
  o.ab_f16 = 1;
  return gemm_impl(probs_f16, heads_keys, vt_block, (int64_t)total_frames * 64, h, dim, bias, h, dim, T, dim, heads_keys, 1,
                   0, o, stream);
}

static int gemm_impl(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
                     const float* resid, int64_t ldr, int M, int N, int K, int out_mode, int bn_hint,
                     const GemmOpts& o, cudaStream_t stream) {
  RALD_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  RALD_REQUIRE(N % 32 == 0, "gemm: N=%d must be a multiple of 32", N);
  RALD_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0, "gemm: K/lda/ldw must be multiples of 8 (16-byte rows)");
  RALD_REQUIRE(out_mode >= 0 && out_mode <= 3, "gemm: out_mode %d", out_mode);
  RALD_REQUIRE(out_mode != 3 || (resid == nullptr && bias == nullptr && N % 64 == 0),
               "gemm: the softmax epilogue takes no bias / residual and needs N a multiple of 64");
  RALD_REQUIRE(o.b_mode == 0 || (!o.w_split && o.grp_rows > 0 && o.grp_rows % (2 * GEMM_BM) == 0 && K % GEMM_BK == 0 &&
                                 N % 64 == 0 && M % o.grp_rows == 0 && o.grp_frame0 >= 0 &&
                                 o.grp_frame0 + M / o.grp_rows <= o.grp_total && (o.b_mode != 2 || K == 8 * 64)),
               "gemm: bad grouped-operand geometry (mode %d, %d rows per frame, frames [%d, +%d) of %d)", o.b_mode,
               o.grp_rows, o.grp_frame0, o.grp_rows > 0 ? M / o.grp_rows : 0, o.grp_total);
  RALD_REQUIRE(out_mode != 2 || resid == nullptr, "gemm: GEGLU epilogue takes no residual");
  RALD_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0 && ldo % (out_mode == 1 ? 4 : 8) == 0,
               "gemm: output not 16-byte aligned");
  RALD_REQUIRE(resid == nullptr || ((reinterpret_cast<uintptr_t>(resid) & 15) == 0 && ldr % 4 == 0),
               "gemm: residual not 16-byte aligned");
  RALD_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm: bias not 16-byte aligned");
  RALD_REQUIRE(o.f16_period == 0 || (o.f16_period % 64 == 0 && o.f16_start % 64 == 0 && o.f16_start < o.f16_period),
               "gemm: fp16 column window start=%d period=%d must be multiples of 64", o.f16_start, o.f16_period);
  RALD_REQUIRE(o.f16_period == 0 || N % 64 == 0, "gemm: N=%d must be a multiple of 64 for mixed fp16 / bf16 output", N);
  RALD_REQUIRE(o.w_k_off % 8 == 0, "gemm: W column shift %d must be a multiple of 8 (TMA box origins are 16-byte aligned)",
               o.w_k_off);
  RALD_REQUIRE(o.tap_n == 0 || (o.split_k && o.n_taps >= 1 && o.n_taps <= 16 && N == o.n_taps * o.tap_n && o.tap_n % 32 == 0),
               "gemm: bad tap layout (%d taps of %d columns, N=%d)", o.n_taps, o.tap_n, N);
  for (int t = 0; t < o.n_taps; ++t)
    RALD_REQUIRE(o.tap_shift[t] % 8 == 0, "gemm: tap shift %d must be a multiple of 8", o.tap_shift[t]);
  RALD_REQUIRE(!o.w_split || K % GEMM_BK == 0, "gemm: split weights need K=%d to be a multiple of %d", K, GEMM_BK);
  const int k_total = o.w_split ? 2 * K : K;   // k extent the main loop walks

  const int sms = device_sm_count();
  const int m_blks = (M + GEMM_BM - 1) / GEMM_BM;
  // narrowest tile of the TMA epilogue (grouped K' operands come in 64-row boxes per head)
  const int min_bn = (out_mode == 1 && o.b_mode == 0) ? 32 : (out_mode == 2 ? 128 : 64);
  int bn = bn_hint;
  if (bn == 0) {
    // Cost model: waves x (fixed per-tile cost + tile width). Wide tiles re-read the operands least and amortise the
    // epilogue, but a small problem wants many narrow tiles so that every SM streams operands (one wave), and a
    // medium one wants the width whose last wave is not mostly empty.
    long best = -1;
    for (int cand = 256; cand >= min_bn; cand >>= 1) {
      if (N % cand != 0) continue;
      const long tiles = (long)m_blks * (N / cand);
      const long cost = ((tiles + sms - 1) / sms) * (64 + cand);
      if (best < 0 || cost < best) { best = cost; bn = cand; }
    }
    if (bn == 0) bn = N % 64 == 0 ? 64 : 32;
  }
  // ---- K splits (weight gradients: K = rows of the batch, few output tiles) ----
  int k_splits = 1, kb_per_split = 0;
  if (o.split_k) {
    RALD_REQUIRE(out_mode == 1 && bias == nullptr && resid == out && ldr == ldo && o.resid_mod == 0 && o.b_mode == 0 &&
                 !o.w_split, "gemm: K splits need the fp32 accumulate-into-out form without bias");
    // widest tile (least re-reads); with taps a tile holds whole taps: tap_n % bn == 0 or bn % tap_n == 0 (tap_n >= 64)
    if (bn_hint == 0) {
      bn = N % 256 == 0 ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 32));
      if (o.tap_n > 0) {
        while (bn > 32 && !((o.tap_n % bn == 0) || (bn % o.tap_n == 0 && o.tap_n % 64 == 0))) bn >>= 1;
      }
    }
    const int num_kb = (K + GEMM_BK - 1) / GEMM_BK;
    const long tiles = (long)m_blks * ((N + bn - 1) / bn);
    int want = (int)(sms / (tiles > 0 ? tiles : 1));
    if (want > 64) want = 64;
    // largest split count <= want that divides the k-blocks into equal groups of a multiple of 4 (ring-slot grouping)
    for (int ks = want; ks >= 2; --ks) {
      if (num_kb % ks == 0 && (num_kb / ks) % 4 == 0) { k_splits = ks; kb_per_split = num_kb / ks; break; }
    }
  }
  RALD_REQUIRE(bn == 32 || bn == 64 || bn == 128 || bn == 256, "gemm: BN=%d unsupported", bn);

  int epi = EPI_GENERIC;
  if (out_mode == 1) {
    if (resid == nullptr) epi = EPI_TMA_STORE;
    else if (resid == out && ldr == ldo && o.resid_mod == 0) epi = EPI_TMA_REDUCE;
  } else if (out_mode == 0 || out_mode == 3) {
    if (resid == nullptr && bn >= 64) epi = EPI_TMA_STORE;
  } else if (bn >= 128) {
    epi = EPI_TMA_STORE;
  }
  RALD_REQUIRE(out_mode != 3 || epi == EPI_TMA_STORE, "gemm: the softmax epilogue needs the TMA-store path (BN >= 64)");
  RALD_REQUIRE(o.b_mode == 0 || (bn >= 64 && epi != EPI_GENERIC), "gemm: grouped operands need BN >= 64 and a TMA epilogue");

  // CTA pairs: 256 x 256 tiles when the problem still fills the machine with them (large-batch regime)
  const int m_blks2 = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const bool pair = gemm_env().pair && bn_hint >= 0 && bn == 256 && epi != EPI_GENERIC && N % 256 == 0 &&
                    (long)m_blks2 * (N / 256) >= sms / 2 && k_splits == 1;

  GemmParams p;
  p.out = out;
  p.ldo = ldo;
  p.bias = bias;
  p.resid = resid;
  p.ldr = ldr;
  p.resid_mod = o.resid_mod;
  p.M = M;
  p.N = N;
  p.K = k_total;
  p.a_k = K;
  p.num_m_blks = pair ? m_blks2 : m_blks;
  p.num_n_blks = (N + bn - 1) / bn;
  p.dbg = g_gemm_dbg;
  p.f16_start = o.f16_start;
  p.f16_period = o.f16_period;
  p.gelu_exact = o.gelu_exact;
  p.b_mode = o.b_mode;
  p.grp_rows = o.grp_rows;
  p.grp_frame0 = o.grp_frame0;
  p.grp_total = o.grp_total;
  p.ab_f16 = o.ab_f16;
  p.w_k_off = o.w_k_off;
  p.panel_len = o.panel_len;
  p.b_halo = o.b_halo;
  p.tap_n = o.tap_n;
  for (int t = 0; t < 16; ++t) p.tap_shift[t] = o.tap_shift[t];
  p.k_splits = k_splits;
  p.kb_per_split = kb_per_split;
  p.w_static = k_splits > 1 ? 0 : (gemm_env().wpre && g_w_static > 0 && !pair && pdl_enabled()) ? 1 : 0;
  RALD_REQUIRE(o.f16_period == 0 || epi == EPI_TMA_STORE, "gemm: mixed fp16 / bf16 output needs the TMA-store epilogue");

  CUtensorMap tmA, tmB, tmO;
  if (o.panel_len > 0) {
    RALD_REQUIRE(o.split_k && o.panel_len % GEMM_BK == 0 && o.n_panels > 0 && K == o.n_panels * o.panel_len &&
                 o.b_halo % 8 == 0 && o.b_mode == 0, "gemm: bad panel layout (%d panels of %d columns, K=%d, halo %d)",
                 o.n_panels, o.panel_len, K, o.b_halo);
    const int w_rows = o.tap_n > 0 ? o.tap_n : N;
    const int w_len = o.panel_len + 2 * o.b_halo;
    uint64_t da[3] = {(uint64_t)o.panel_len, (uint64_t)M, (uint64_t)o.n_panels};
    uint64_t sa[2] = {(uint64_t)o.panel_len * 2, (uint64_t)M * o.panel_len * 2};
    uint32_t ba[3] = {64, (uint32_t)GEMM_BM, 1};
    RALD_TRY(make_tmap_nd_bf16(&tmA, A, 3, da, sa, ba, nullptr));
    uint64_t db[3] = {(uint64_t)w_len, (uint64_t)w_rows, (uint64_t)o.n_panels};
    uint64_t sb[2] = {(uint64_t)w_len * 2, (uint64_t)w_rows * w_len * 2};
    uint32_t bb[3] = {64, (uint32_t)((o.tap_n > 0 && o.tap_n < bn) ? o.tap_n : bn), 1};
    RALD_TRY(make_tmap_nd_bf16(&tmB, W, 3, db, sb, bb, nullptr));
  } else {
  RALD_TRY(make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, (uint32_t)GEMM_BM));
  if (o.b_mode == 1) {          // K' of one block: [8 heads * grp_total frames * 64 keys][K], one 64-row box per head
    RALD_TRY(make_tmap_2d_bf16(&tmB, W, (uint64_t)(N / 64) * o.grp_total * 64, (uint64_t)K, (uint64_t)ldw, 64u));
  } else if (o.b_mode == 2) {   // VT of one block: [8 heads * N][grp_total * 64]
    RALD_TRY(make_tmap_2d_bf16(&tmB, W, (uint64_t)(K / 64) * N, (uint64_t)o.grp_total * 64, (uint64_t)ldw,
                               (uint32_t)(pair ? bn / 2 : bn)));
  } else {
    RALD_TRY(make_tmap_2d_bf16(&tmB, W, (uint64_t)(o.tap_n > 0 ? o.tap_n : N), (uint64_t)k_total, (uint64_t)ldw,
                               (uint32_t)(pair ? bn / 2 : ((o.tap_n > 0 && o.tap_n < bn) ? o.tap_n : bn))));
  }
  }
  if (epi != EPI_GENERIC) {
    RALD_TRY(make_tmap_out(&tmO, out, (uint64_t)M, (uint64_t)(out_mode == 2 ? N / 2 : N), (uint64_t)ldo, out_mode == 1, 32u));
  } else {
    tmO = tmA;
  }

  if (pair) {
    if (out_mode == 3) return launch_gemm<256, 3, EPI_TMA_STORE, 2>(tmA, tmB, tmO, p, sms, stream);
    if (out_mode == 0) return launch_gemm<256, 0, EPI_TMA_STORE, 2>(tmA, tmB, tmO, p, sms, stream);
    if (out_mode == 2) return launch_gemm<256, 2, EPI_TMA_STORE, 2>(tmA, tmB, tmO, p, sms, stream);
    if (epi == EPI_TMA_REDUCE) return launch_gemm<256, 1, EPI_TMA_REDUCE, 2>(tmA, tmB, tmO, p, sms, stream);
    return launch_gemm<256, 1, EPI_TMA_STORE, 2>(tmA, tmB, tmO, p, sms, stream);
  }
#define RALD_GEMM_LAUNCH(BN_, MODE_, EPI_) return launch_gemm<BN_, MODE_, EPI_>(tmA, tmB, tmO, p, sms, stream)
#define RALD_GEMM_BN(MODE_, EPI_)                   \
  switch (bn) {                                     \
    case 32: RALD_GEMM_LAUNCH(32, MODE_, EPI_);     \
    case 64: RALD_GEMM_LAUNCH(64, MODE_, EPI_);     \
    case 128: RALD_GEMM_LAUNCH(128, MODE_, EPI_);   \
    default: RALD_GEMM_LAUNCH(256, MODE_, EPI_);    \
  }
  if (out_mode == 1) {
    if (epi == EPI_TMA_REDUCE) { RALD_GEMM_BN(1, EPI_TMA_REDUCE) }
    if (epi == EPI_TMA_STORE) { RALD_GEMM_BN(1, EPI_TMA_STORE) }
    RALD_GEMM_BN(1, EPI_GENERIC)
  }
  if (out_mode == 3) {
    switch (bn) {
      case 64: RALD_GEMM_LAUNCH(64, 3, EPI_TMA_STORE);
      case 128: RALD_GEMM_LAUNCH(128, 3, EPI_TMA_STORE);
      default: RALD_GEMM_LAUNCH(256, 3, EPI_TMA_STORE);
    }
  }
  if (out_mode == 0) {
    if (epi == EPI_TMA_STORE) {
      switch (bn) {
        case 64: RALD_GEMM_LAUNCH(64, 0, EPI_TMA_STORE);
        case 128: RALD_GEMM_LAUNCH(128, 0, EPI_TMA_STORE);
        default: RALD_GEMM_LAUNCH(256, 0, EPI_TMA_STORE);
      }
    }
    RALD_GEMM_BN(0, EPI_GENERIC)
  }
  if (epi == EPI_TMA_STORE) {
    if (bn == 128) RALD_GEMM_LAUNCH(128, 2, EPI_TMA_STORE);
    RALD_GEMM_LAUNCH(256, 2, EPI_TMA_STORE);
  }
  RALD_GEMM_BN(2, EPI_GENERIC)
#undef RALD_GEMM_BN
#undef RALD_GEMM_LAUNCH
  return -1;
}

}  // namespace rald

// Debug hook (tools/gemm_phases.py): when set, every GEMM CTA stores %globaltimer at its phase boundaries.
extern "C" int rald_gemm_debug_buffer(unsigned long long* dev_buf) {
  rald::g_gemm_dbg = dev_buf;
  return 0;
}
