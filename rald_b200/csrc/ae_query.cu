// Streaming decoder-query kernel of the VecSet autoencoder (KLAutoEncoder.decode, model/models_ae.py:417-424):
//   logit(q) = to_outputs( to_out( softmax( to_q(LN(point_embed(q))) K^T / sqrt(d) ) V ) )
// for up to millions of query points per frame against that frame's 512-latent context.
//
// Exact algebraic folding (no residual / feed-forward follows the decoder cross-attention, decoder_ff=False,
// and to_outputs is Linear(dim, 1)):   sim = LN(e) (K W_q)^T / sqrt(d)   and   logit = softmax(sim) . v' + c0
// with K' = K W_q  [512 latents x 512],  v' = V W_out^T w_o^T [512],  c0 = b_out . w_o + b_o   (SURVEY.md §7.3).
// K' (bf16), v' (fp32) and c0 are per-frame constants prepared once per latent set by the host runtime.
//
// One CTA = a persistent loop over 128-query tiles; per tile, entirely on chip:
//   compute warps : Fourier features [sin(p f), cos(p f), p] -> bf16 A tile in 128B-swizzled smem
//   MMA (tcgen05) : E = feat x W_pe^T                      (M=128, N=512, K=64)       -> TMEM cols [0,512)
//   compute warps : e + bias, LayerNorm (fp32, two sweeps over TMEM) -> bf16 pairs written back over E
//                   as the TMEM-resident A operand Qn (k-blocks 0..3 in cols [0,128), 4..7 in cols [256,384))
//   MMA (tcgen05) : S_i = Qn x K'_i^T for 4 chunks of 128 latents (A from TMEM, K' streamed by TMA through a
//                   smem ring), double-buffered in the freed TMEM cols [128,256) / [384,512)
//   compute warps : online softmax over the chunks and the dot with v' in fp32 registers -> one logit per query.
// Eight compute warps: two per TMEM lane quarter, each owning one half of the columns (of E for the LayerNorm, of
// every S chunk for the softmax); the halves' (sum, sum of squares) and (max, sum, dot) are merged exactly through
// shared memory. All TMEM loads are software-pipelined (the next 32 columns are in flight during the arithmetic).
// Mandatory HBM traffic is 12 B in + 4 B out per query; the reference materialises ~8 [B,Q,512] fp32 tensors.
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int AQ_TILE = 128;          // queries per tile
constexpr int AQ_DIM = 512;           // model width = latents per frame (both 512 in every reference config)
constexpr int AQ_FEAT = 64;           // 51 Fourier features padded to one 128-byte bf16 row
constexpr int AQ_STAGES = 8;          // smem ring of [128 rows x 64 cols] bf16 B tiles (16 KB each)
constexpr int AQ_STAGE_BYTES = 128 * 64 * 2;
constexpr int AQ_THREADS = 320;       // TMA warp, MMA warp, 8 compute warps
constexpr int AQ_CWARPS = 8;
// AQ_CLUSTER > 1: the CTAs of a cluster work on neighbouring query tiles of the SAME frame in lockstep and share
// every operand tile of the ring — each CTA fetches 1 / AQ_CLUSTER of a tile's rows and multicasts them to all (TMA
// multicast), a slot is refilled once the MMAs of all CTAs have released it (multicast commit). Measured on B200 with
// AQ_CLUSTER = 2: parity-green but NO gain (2.65 vs 2.7 us per 128-latent chunk): the chunk time is set by the 32
// tcgen05.mma with the A operand in TMEM (~85 ns each), not by operand delivery from L2. Left at 1.
constexpr int AQ_CLUSTER = 1;

struct AeQueryParams {
  const float* queries;   // [B, Q, 3]
  float* logits;          // [B, Q]
  const float* pe_bias;   // [512]
  const float* ln_g;      // [512] decoder_cross_attn.norm.weight
  const float* ln_b;      // [512]
  const float* vprime;    // [B, 512]
  const float* c0;        // [B]
  float freq[3][8];       // point_embed.basis block diagonal
  int octaves;            // 1: freq[a][k] == freq[a][0] * 2^k exactly (the constructor's pi * 2^k): double-angle path
  int B;
  int64_t Q;
  int tiles_per_frame;    // padded to a multiple of AQ_CLUSTER (tiles beyond Q compute nothing that is stored)
  float scale_log2;       // log2(e) / sqrt(dim)
  unsigned long long* dbg;  // optional [tile < 4][16] %globaltimer stamps of CTA 0's first compute warp
};

static unsigned long long* g_aq_dbg = nullptr;
#define AQ_STAMP(slot_)                                                                                       \
  do {                                                                                                        \
    if (p.dbg != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0 && tile_n < 4)                          \
      p.dbg[tile_n * 16 + (slot_)] = global_timer_ns();                                                       \
  } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AQ_THREADS, 1)
ae_query_kernel(const __grid_constant__ CUtensorMap tmWpe, const __grid_constant__ CUtensorMap tmKp,
                const AeQueryParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_feat = smem;                                  // 16 KB
  uint8_t* s_ring = s_feat + AQ_TILE * AQ_FEAT * 2;        // AQ_STAGES x 16 KB
  float* s_bias = reinterpret_cast<float*>(s_ring + AQ_STAGES * AQ_STAGE_BYTES);
  float* s_g = s_bias + AQ_DIM;
  float* s_b = s_g + AQ_DIM;
  float* s_v = s_b + AQ_DIM;                               // v' of the current frame
  float* s_x = s_v + AQ_DIM;                               // [2 halves][3][128] exchange between the column halves
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + 2 * 3 * AQ_TILE);
  uint64_t* full_bar = bars;                    // [AQ_STAGES]
  uint64_t* empty_bar = bars + AQ_STAGES;       // [AQ_STAGES]
  uint64_t* feat_ready = bars + 2 * AQ_STAGES;  // compute -> MMA
  uint64_t* e_ready = feat_ready + 1;           // MMA -> compute
  uint64_t* qn_ready = e_ready + 1;             // compute -> MMA
  uint64_t* s_ready = qn_ready + 1;             // [2] MMA -> compute
  uint64_t* s_free = s_ready + 2;               // [2] compute -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.B * p.tiles_per_frame;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmWpe);
    tma_prefetch_desc(&tmKp);
    for (int s = 0; s < AQ_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], AQ_CLUSTER);   // one multicast commit from every CTA of the cluster
    }
    mbar_init(feat_ready, AQ_CWARPS);
    mbar_init(e_ready, 1);
    mbar_init(qn_ready, AQ_CWARPS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&s_free[i], AQ_CWARPS);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < AQ_DIM; i += blockDim.x) {
    s_bias[i] = p.pe_bias[i];
    s_g[i] = p.ln_g[i];
    s_b[i] = p.ln_b[i];
  }
  tc_fence_before();
  if (AQ_CLUSTER > 1) cluster_sync_all();   // every CTA's barriers are initialised before a multicast can reach them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = cluster_ctarank();
  constexpr uint16_t CMASK = (1u << AQ_CLUSTER) - 1u;
  constexpr int SLICE_ROWS = 128 / AQ_CLUSTER;
  constexpr int SLICE_BYTES = AQ_STAGE_BYTES / AQ_CLUSTER;

  if (warp == 0) {
    // ===================== TMA producer: 4 W_pe tiles then 4 x 8 K' tiles per query tile =====================
    // (this CTA's row slice of every tile, multicast to the whole cluster; the tile is complete in a CTA when all
    // AQ_CLUSTER slices have landed = AQ_STAGE_BYTES on its own full barrier)
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int frame = tile / p.tiles_per_frame;
        for (int n = 0; n < 4; ++n) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], AQ_STAGE_BYTES);
          if (AQ_CLUSTER == 1) tma_load_2d(s_ring + s * AQ_STAGE_BYTES, &tmWpe, &full_bar[s], 0, n * 128);
          else tma_load_2d_mcast(s_ring + s * AQ_STAGE_BYTES + crank * SLICE_BYTES, &tmWpe, &full_bar[s], 0,
                                 n * 128 + (int)crank * SLICE_ROWS, CMASK);
          if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
        }
        for (int i = 0; i < 4; ++i) {
          for (int kb = 0; kb < 8; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], AQ_STAGE_BYTES);
            if (AQ_CLUSTER == 1)
              tma_load_2d(s_ring + s * AQ_STAGE_BYTES, &tmKp, &full_bar[s], kb * 64, frame * AQ_DIM + i * 128);
            else
              tma_load_2d_mcast(s_ring + s * AQ_STAGE_BYTES + crank * SLICE_BYTES, &tmKp, &full_bar[s], kb * 64,
                                frame * AQ_DIM + i * 128 + (int)crank * SLICE_ROWS, CMASK);
            if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      constexpr uint32_t idesc = make_idesc(FMT_BF16, 128, 128, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      uint32_t tile_ph = 0;       // parity of feat_ready / qn_ready (one completion per tile)
      uint32_t free_ph[2] = {0, 0};
      const uint64_t a_desc = make_sdesc_sw128(smem_u32(s_feat), 16, 1024);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // ---- E = feat x W_pe^T ----
        mbar_wait(feat_ready, tile_ph);
        tc_fence_after();
        for (int n = 0; n < 4; ++n) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t b_desc = make_sdesc_sw128(smem_u32(s_ring + s * AQ_STAGE_BYTES), 16, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) mma_f16_ss(tmem_base + n * 128, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
          if (AQ_CLUSTER == 1) tc_commit(&empty_bar[s]);
          else tc_commit_mcast(&empty_bar[s], CMASK);
          if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(e_ready);
        // ---- S_i = Qn x K'_i^T, Qn read from TMEM cols [0,128) (k-blocks 0..3) and [256,384) (k-blocks 4..7) ----
        mbar_wait(qn_ready, tile_ph);
        tc_fence_after();
        for (int i = 0; i < 4; ++i) {
          const int buf = i & 1;
          mbar_wait(&s_free[buf], free_ph[buf] ^ 1);
          free_ph[buf] ^= 1;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + 128 + buf * 256;
          for (int kb = 0; kb < 8; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint64_t b_desc = make_sdesc_sw128(smem_u32(s_ring + s * AQ_STAGE_BYTES), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_ts(d_tmem, tmem_base + (kb < 4 ? kb * 32 : 256 + (kb - 4) * 32) + k * 8, b_desc + 2 * k, idesc,
                         (kb | k) != 0);
            if (AQ_CLUSTER == 1) tc_commit(&empty_bar[s]);
            else tc_commit_mcast(&empty_bar[s], CMASK);
            if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
          }
          tc_commit(&s_ready[buf]);
        }
        tile_ph ^= 1;
      }
    }
  } else {
    // ===================== compute warps (256 threads: thread pair <-> query row <-> TMEM lane) =====================
    const int q4 = warp & 3;               // TMEM lane quarter
    const int hs = (warp - 2) >> 2;        // column half owned by this warp
    const int r = q4 * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    const int ct = threadIdx.x - 64;       // 0..255 among the compute threads
    float* x_mine = s_x + hs * 3 * AQ_TILE;
    const float* x_other = s_x + (hs ^ 1) * 3 * AQ_TILE;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(2 + q4) : "memory"); };  // the row's two warps
    uint32_t tile_ph = 0;
    uint32_t ready_ph[2] = {0, 0};
    int cur_frame = -1;
    int tile_n = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ++tile_n;
      AQ_STAMP(0);
      const int frame = tile / p.tiles_per_frame;
      const int64_t q0 = static_cast<int64_t>(tile - frame * p.tiles_per_frame) * AQ_TILE;
      const int64_t qi = q0 + r;
      const bool q_ok = qi < p.Q;
      if (frame != cur_frame) {
        // all compute threads finished the previous tile's reads of s_v before it is replaced
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = ct; i < AQ_DIM; i += 32 * AQ_CWARPS) s_v[i] = p.vprime[(int64_t)frame * AQ_DIM + i];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        cur_frame = frame;
      }
      // ---- Fourier features (models_ae.py:128-137): [sin(p f) (24), cos(p f) (24), p (3)], zero padded to 64 ----
      // 16-byte chunk c of the 128-byte row holds features 8c..8c+7. Warp half 0: frequencies 0..15 of the 24
      // (x: 8, y: 8) -> chunks 0, 1 (sin) and 3, 4 (cos); half 1: frequencies 16..23 (z) -> chunks 2, 5, and 6, 7
      // (the raw point and the padding).
      float pt[3] = {0.f, 0.f, 0.f};
      if (q_ok) {
        const float* qp = p.queries + ((int64_t)frame * p.Q + qi) * 3;
        pt[0] = qp[0]; pt[1] = qp[1]; pt[2] = qp[2];
      }
      {
        uint8_t* row = s_feat + r * 128;
        auto put = [&](int c, const float (&f)[8]) {
          const uint4 v = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                                     pack_bf16x2(f[6], f[7]));
          *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = v;  // 128-byte swizzle: chunk ^= row % 8
        };
        const int a0 = hs == 0 ? 0 : 2, a1 = hs == 0 ? 2 : 3;
        for (int a = a0; a < a1; ++a) {
          float sn[8], cs[8];
          if (p.octaves) {
            // sin / cos of x, 2x, 4x, ... by angle doubling from ONE accurate sincos: the absolute error doubles per
            // octave (<= 2^7 x 1e-7 at the top one) — the same size as the reference's own fp32 rounding of the
            // argument p * pi * 2^k (|arg| up to 402), and far below the bf16 rounding of the features (4e-3)
            sincosf(pt[a] * p.freq[a][0], &sn[0], &cs[0]);
#pragma unroll
            for (int k = 1; k < 8; ++k) {
              sn[k] = 2.0f * sn[k - 1] * cs[k - 1];
              cs[k] = fmaf(-2.0f * sn[k - 1], sn[k - 1], 1.0f);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) sincosf(pt[a] * p.freq[a][k], &sn[k], &cs[k]);
          }
          put(a, sn);
          put(3 + a, cs);
        }
        if (hs == 1) {
          const float tail[8] = {pt[0], pt[1], pt[2], 0.f, 0.f, 0.f, 0.f, 0.f};
          const float zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          put(6, tail);
          put(7, zero);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(feat_ready);
      AQ_STAMP(1);

      // ---- LayerNorm of e = E + bias (fp32); this warp owns columns [256 hs, 256 hs + 256) ----
      mbar_wait(e_ready, tile_ph);
      tc_fence_after();
      AQ_STAMP(2);
      const int c0 = 256 * hs;
      float sum = 0.f, sq = 0.f;
      {
        uint32_t va[32], vb[32];
        auto acc32 = [&](const uint32_t (&v)[32], int c) {
          // 128-bit shared-memory reads: the per-column constants are warp-wide broadcasts, and one LDS per value
          // made the load/store pipe the limiter of both sweeps
          float s0 = 0.f, s1 = 0.f, q0_ = 0.f, q1 = 0.f;
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j];
            const float e0 = __uint_as_float(v[4 * j]) + bb.x, e1 = __uint_as_float(v[4 * j + 1]) + bb.y;
            const float e2 = __uint_as_float(v[4 * j + 2]) + bb.z, e3 = __uint_as_float(v[4 * j + 3]) + bb.w;
            s0 += e0 + e2; s1 += e1 + e3;
            q0_ = fmaf(e0, e0, q0_); q1 = fmaf(e1, e1, q1);
            q0_ = fmaf(e2, e2, q0_); q1 = fmaf(e3, e3, q1);
          }
          sum += s0 + s1;
          sq += q0_ + q1;
        };
        tmem_ld32(t_lane + c0, va);
#pragma unroll 1
        for (int c = 0; c < 256; c += 64) {
          tmem_ld_wait();
          tmem_ld32(t_lane + c0 + c + 32, vb);
          acc32(va, c0 + c);
          tmem_ld_wait();
          if (c + 64 < 256) tmem_ld32(t_lane + c0 + c + 64, va);
          acc32(vb, c0 + c + 32);
        }
      }
      x_mine[r] = sum;
      x_mine[AQ_TILE + r] = sq;
      AQ_STAMP(3);
      pair_sync();
      sum += x_other[r];
      sq += x_other[AQ_TILE + r];
      const float mean = sum * (1.0f / AQ_DIM);
      const float var = fmaxf(sq * (1.0f / AQ_DIM) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      {
        // Qn (bf16 pairs) of columns [c0 + c, c0 + c + 32) -> TMEM columns [c0 + c/2, +16): inside this warp's own,
        // already consumed E columns (k-blocks 0..3 end up in [0,128), k-blocks 4..7 in [256,384))
        uint32_t va[32], vb[32];
        auto norm32 = [&](const uint32_t (&v)[32], int c) {
          uint32_t pk[16];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + c);
          const float4* g4 = reinterpret_cast<const float4*>(s_g + c);
          const float4* o4 = reinterpret_cast<const float4*>(s_b + c);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[j], gg = g4[j], oo = o4[j];
            const float e0 = (__uint_as_float(v[4 * j]) + bb.x - mean) * rstd;
            const float e1 = (__uint_as_float(v[4 * j + 1]) + bb.y - mean) * rstd;
            const float e2 = (__uint_as_float(v[4 * j + 2]) + bb.z - mean) * rstd;
            const float e3 = (__uint_as_float(v[4 * j + 3]) + bb.w - mean) * rstd;
            pk[2 * j] = pack_bf16x2(fmaf(e0, gg.x, oo.x), fmaf(e1, gg.y, oo.y));
            pk[2 * j + 1] = pack_bf16x2(fmaf(e2, gg.z, oo.z), fmaf(e3, gg.w, oo.w));
          }
          tmem_st16(t_lane + c0 + ((c - c0) >> 1), pk);
        };
        tmem_ld32(t_lane + c0, va);
#pragma unroll 1
        for (int c = 0; c < 256; c += 64) {
          tmem_ld_wait();
          tmem_ld32(t_lane + c0 + c + 32, vb);     // columns beyond every store issued so far
          norm32(va, c0 + c);
          tmem_ld_wait();
          if (c + 64 < 256) tmem_ld32(t_lane + c0 + c + 64, va);
          norm32(vb, c0 + c + 32);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qn_ready);
      AQ_STAMP(4);

      // ---- online softmax over 4 x 128 latents, fused with the dot against v'; this warp owns 64 latents of each ----
      float m = -INFINITY, l = 0.f, acc = 0.f;
      for (int i = 0; i < 4; ++i) {
        const int buf = i & 1;
        mbar_wait(&s_ready[buf], ready_ph[buf]);
        ready_ph[buf] ^= 1;
        tc_fence_after();
        AQ_STAMP(5 + i);
        const uint32_t t_s = t_lane + 128 + buf * 256 + 64 * hs;
        uint32_t va[32], vb[32];
        tmem_ld32(t_s, va);
        tmem_ld32(t_s + 32, vb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[buf]);   // both halves are in registers: the buffer may be overwritten
        float cm0 = __uint_as_float(va[0]), cm1 = __uint_as_float(vb[0]);
#pragma unroll
        for (int j = 1; j < 32; ++j) {
          cm0 = fmaxf(cm0, __uint_as_float(va[j]));
          cm1 = fmaxf(cm1, __uint_as_float(vb[j]));
        }
        const float m_new = fmaxf(m, fmaxf(cm0, cm1) * p.scale_log2);
        const float corr = ex2f(m - m_new);  // first group: ex2(-inf) = 0
        m = m_new;
        const float* vp = s_v + i * 128 + 64 * hs;
        float l0 = 0.f, l1 = 0.f, a0 = 0.f, a1 = 0.f;
        const float4* vp4 = reinterpret_cast<const float4*>(vp);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 wa = vp4[j], wb = vp4[8 + j];
          const float pa0 = ex2f(fmaf(__uint_as_float(va[4 * j]), p.scale_log2, -m));
          const float pa1 = ex2f(fmaf(__uint_as_float(va[4 * j + 1]), p.scale_log2, -m));
          const float pa2 = ex2f(fmaf(__uint_as_float(va[4 * j + 2]), p.scale_log2, -m));
          const float pa3 = ex2f(fmaf(__uint_as_float(va[4 * j + 3]), p.scale_log2, -m));
          const float pb0 = ex2f(fmaf(__uint_as_float(vb[4 * j]), p.scale_log2, -m));
          const float pb1 = ex2f(fmaf(__uint_as_float(vb[4 * j + 1]), p.scale_log2, -m));
          const float pb2 = ex2f(fmaf(__uint_as_float(vb[4 * j + 2]), p.scale_log2, -m));
          const float pb3 = ex2f(fmaf(__uint_as_float(vb[4 * j + 3]), p.scale_log2, -m));
          l0 += (pa0 + pa1) + (pa2 + pa3);
          l1 += (pb0 + pb1) + (pb2 + pb3);
          a0 = fmaf(pa0, wa.x, a0); a0 = fmaf(pa1, wa.y, a0); a0 = fmaf(pa2, wa.z, a0); a0 = fmaf(pa3, wa.w, a0);
          a1 = fmaf(pb0, wb.x, a1); a1 = fmaf(pb1, wb.y, a1); a1 = fmaf(pb2, wb.z, a1); a1 = fmaf(pb3, wb.w, a1);
        }
        l = fmaf(l, corr, l0 + l1);
        acc = fmaf(acc, corr, a0 + a1);
      }
      AQ_STAMP(9);
      // ---- merge the two column halves (exact): half 1 hands (m, l, acc) to half 0 ----
      if (hs == 1) { x_mine[r] = m; x_mine[AQ_TILE + r] = l; x_mine[2 * AQ_TILE + r] = acc; }
      pair_sync();
      if (hs == 0) {
        const float m1 = x_other[r], l1 = x_other[AQ_TILE + r], acc1 = x_other[2 * AQ_TILE + r];
        const float mm = fmaxf(m, m1);
        const float w0 = ex2f(m - mm), w1 = ex2f(m1 - mm);
        if (q_ok) p.logits[(int64_t)frame * p.Q + qi] = (acc * w0 + acc1 * w1) / (l * w0 + l1 * w1) + p.c0[frame];
      }
      tile_ph ^= 1;
    }
  }

  tc_fence_before();
  if (AQ_CLUSTER > 1) cluster_sync_all();   // no CTA exits while a peer may still multicast into its shared memory
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int ae_query(const float* queries, int B, int64_t Q, const void* wpe_bf16, const float* pe_bias, const float* ln_g,
             const float* ln_b, const void* kprime_bf16, const float* vprime, const float* c0, const float* freq24,
             float* logits, int dim, int n_latents, cudaStream_t stream) {
  RALD_REQUIRE(dim == AQ_DIM && n_latents == AQ_DIM, "ae_query: dim=%d latents=%d unsupported (512/512 only)", dim,
               n_latents);
  RALD_REQUIRE(B > 0 && Q > 0, "ae_query: empty batch");
  CUtensorMap tmWpe, tmKp;
  RALD_TRY(make_tmap_2d_bf16(&tmWpe, wpe_bf16, AQ_DIM, AQ_FEAT, AQ_FEAT, 128 / AQ_CLUSTER));
  RALD_TRY(make_tmap_2d_bf16(&tmKp, kprime_bf16, (uint64_t)B * AQ_DIM, AQ_DIM, AQ_DIM, 128 / AQ_CLUSTER));
  AeQueryParams p;
  p.queries = queries; p.logits = logits; p.pe_bias = pe_bias; p.ln_g = ln_g; p.ln_b = ln_b; p.vprime = vprime;
  p.c0 = c0;
  for (int a = 0; a < 3; ++a)
    for (int k = 0; k < 8; ++k) p.freq[a][k] = freq24[a * 8 + k];
  p.octaves = 1;
  for (int a = 0; a < 3; ++a)
    for (int k = 1; k < 8; ++k)
      if (p.freq[a][k] != p.freq[a][0] * (float)(1 << k)) p.octaves = 0;
  p.B = B; p.Q = Q;
  p.tiles_per_frame = (int)((Q + AQ_TILE - 1) / AQ_TILE);
  p.tiles_per_frame = (p.tiles_per_frame + AQ_CLUSTER - 1) / AQ_CLUSTER * AQ_CLUSTER;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)dim);
  p.dbg = g_aq_dbg;
  const int smem_bytes = AQ_TILE * AQ_FEAT * 2 + AQ_STAGES * AQ_STAGE_BYTES + 4 * AQ_DIM * 4 + 2 * 3 * AQ_TILE * 4 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(ae_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int64_t tiles = (int64_t)B * p.tiles_per_frame;
  const int sms = device_sm_count();
  // whole clusters only; every CTA of a cluster runs the same number of tiles (tiles is a multiple of AQ_CLUSTER and
  // the grid-stride keeps cluster mates on neighbouring tiles of one frame)
  const int grid = (int)(tiles < sms ? tiles : sms) / AQ_CLUSTER * AQ_CLUSTER;
  ProfScope prof(FAM_AE_QUERY, stream, (double)B * Q * 577536.0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(AQ_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = AQ_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RALD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ae_query_kernel, tmWpe, tmKp, p));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

// Debug hook like rald_gemm_debug_buffer: the first compute warp of CTA 0 stores %globaltimer stamps of its first 4
// tiles at dev_buf[tile*16 + i]: 0 start, 1 features written, 2 E ready, 3 row statistics, 4 Qn written, 5..8 S chunk
// ready, 9 softmax done.
extern "C" int rald_ae_query_debug_buffer(unsigned long long* dev_buf) {
  rald::g_aq_dbg = dev_buf;
  return 0;
}
