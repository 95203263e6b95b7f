// Fused softmax(Q K^T * scale) V for head_dim 64 and a key set that fits TMEM (Skv <= 512), tcgen05 + TMA.
//
// Persistent CTAs (one per SM). A work item = TQ consecutive 128-query tiles of one (frame, head): K and V of the
// head are loaded ONCE per item and stay in shared memory while the query tiles stream through a 2-deep Q ring.
// The keys are split in two halves that are processed CONCURRENTLY by two softmax warp groups and merged at the end
// (exact: each half keeps its own max / sum, the epilogue rescales):
//   warp 0 lane 0 : TMA producer   (K, V per item; Q per tile, prefetched one tile ahead)
//   warp 1 lane 0 : tcgen05 issuer  S_h = Q K_h^T  (SS, fp32 S in TMEM), O_h = P_h V_h (TS: P from TMEM, V MN-major)
//   warps 2..5    : softmax of key half 0      } thread <-> query row <-> TMEM lane; ONE sweep over S with a lazily
//   warps 6..9    : softmax of key half 1      } raised shift: p = exp2(s*scale*log2e - m) packed to bf16 pairs written
//                   back INTO TMEM over the consumed S columns, row sum in fp32 (TMEM -> register bandwidth, 64 B/clk
//                   per SM, is what bounds this kernel, so S is read exactly once). TMEM loads are software-pipelined
//                   and the reductions use four independent chains. Then both groups exchange (shift, sum) through
//                   shared memory and each normalises and stores 32 of the 64 output columns:
//                   O = (w0 O_0 + w1 O_1) / (w0 l_0 + w1 l_1),  w_h = exp2(m_h - max(m_0, m_1)).
// With 8 softmax warps every SM sub-partition has two warps to hide TMEM-load and MUFU latency; the tensor pipe works
// on one half while the other is in its softmax. For short contexts (Skv <= 128, the cross-attention case) the TMEM
// regions are double-buffered so consecutive query tiles overlap as well. The score matrix never leaves the SM (the
// reference materialises [B*h, Sq, Skv] fp32 in HBM: model/models_radar_generation.py:66-75, models_ae.py:91-104).
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int ATT_BM = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 320;
constexpr int ATT_QBYTES = ATT_BM * ATT_D * 2;

struct AttnParams {
  __nv_bfloat16* out;
  int64_t ldo;
  int Sq, Skv;
  int frames, heads;
  int q_tiles;         // Sq / 128
  int tq;              // query tiles per work item
  int items_per_head;  // q_tiles / tq
  int num_items;
  float scale_log2;    // scale * log2(e)
  int half;            // keys per part = Skv / parts
  int parts;           // 2: the two softmax groups each take one key half. 4 (Skv = 512): key quarters, two passes —
                       // the P V products of the first two quarters run while the groups are in their second pass
  uint32_t o_col;      // O accumulator column inside a half region
  uint32_t region;     // TMEM columns per half region
  int nbuf;            // query tiles in flight in TMEM (1 or 2); a tile uses 2 * region columns
  int kv_frame_rows;   // rows per frame in the K / V tensors (>= Skv; a key CHUNK of a longer context when larger)
  float* stats;        // optional [frames*Sq][heads][2] = (shift m in log2 units, sum l): softmax statistics of the
                       // keys seen by this call, for merging key chunks of a long context (attn_merge_chunks)
  unsigned long long* dbg;  // optional [tile < 16][group 2][8] %globaltimer stamps of CTA 0 (tools/attn_phases.py)
};

static unsigned long long* g_attn_dbg = nullptr;
#define ATT_STAMP(slot_)                                                                              \
  do {                                                                                                \
    if (p.dbg != nullptr && blockIdx.x == 0 && qn < 16 && q == 0 && lane == 0)                        \
      p.dbg[(qn * 2 + g) * 8 + (slot_)] = global_timer_ns();                                          \
  } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_d64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Skv = p.Skv;
  uint8_t* sQ = smem;                       // [2][128 x 64]
  uint8_t* sK = sQ + 2 * ATT_QBYTES;        // [Skv x 64]
  uint8_t* sV = sK + Skv * ATT_D * 2;       // [Skv x 64]
  uint8_t* sO = sV + Skv * ATT_D * 2;       // [128 x 64] bf16 output staging tile (128-byte swizzled rows) for the TMA store
  float* s_stat = reinterpret_cast<float*>(sO + ATT_QBYTES);  // [nbuf 2][half 2][m, l][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 2 * 2 * 2 * ATT_BM);
  uint64_t* k_full = bars + 0;
  uint64_t* k_empty = bars + 1;
  uint64_t* q_full = bars + 2;    // [2]
  uint64_t* q_empty = bars + 4;   // [2]
  uint64_t* s_full = bars + 6;    // [slot 2][half 2]
  uint64_t* p_full = bars + 10;   // [slot 2][half 2]
  uint64_t* o_full = bars + 14;   // [slot 2][half 2]
  uint64_t* o_empty = bars + 18;  // [slot 2]
  uint64_t* v_full = bars + 20;
  uint64_t* v_empty = bars + 21;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&o_empty[i], 8);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
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
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      uint32_t kv_ph = 0;
      int qn = 0;  // running tile counter -> Q ring slot / phase
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        const int fh = item / p.items_per_head;
        const int frame = fh / p.heads, head = fh - frame * p.heads;
        const int qt0 = (item - fh * p.items_per_head) * p.tq;
        const int kv_row0 = frame * p.kv_frame_rows;
        // K is free as soon as the last S product of the previous item has retired, V only after its last P V:
        // the next item's K and first Q tile are fetched while the previous item is still in its softmax.
        mbar_wait(k_empty, kv_ph ^ 1);
        mbar_arrive_expect_tx(k_full, Skv * ATT_D * 2);
        for (int r = 0; r < Skv; r += p.half) tma_load_2d(sK + r * ATT_D * 2, &tmK, k_full, head * ATT_D, kv_row0 + r);
        for (int t = 0; t < p.tq; ++t, ++qn) {
          const int slot = qn & 1;
          const uint32_t ph = (qn >> 1) & 1;
          mbar_wait(&q_empty[slot], ph ^ 1);
          mbar_arrive_expect_tx(&q_full[slot], ATT_QBYTES);
          tma_load_2d(sQ + slot * ATT_QBYTES, &tmQ, &q_full[slot], head * ATT_D, frame * p.Sq + (qt0 + t) * ATT_BM);
          if (t == 0) {
            mbar_wait(v_empty, kv_ph ^ 1);
            mbar_arrive_expect_tx(v_full, Skv * ATT_D * 2);
            for (int r = 0; r < Skv; r += p.half)
              tma_load_2d(sV + r * ATT_D * 2, &tmV, v_full, head * ATT_D, kv_row0 + r);
          }
        }
        kv_ph ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      const uint32_t idesc_o = make_idesc(FMT_F16, ATT_BM, ATT_D, 0, 1);  // P fp16 (TMEM) x V fp16 (smem)
      const uint32_t idesc_s = make_idesc(FMT_BF16, ATT_BM, p.half, 0, 0);
      uint32_t kv_ph = 0;
      int qn = 0;
      auto issue_s = [&](int n, bool last_of_item) {
        const int qslot = n & 1;
        const int slot = p.nbuf == 2 ? (n & 1) : 0;
        const uint32_t use = p.nbuf == 2 ? (uint32_t)(n >> 1) : (uint32_t)n;  // earlier uses of this TMEM slot
        mbar_wait(&q_full[qslot], (n >> 1) & 1);
        mbar_wait(&o_empty[slot], (use & 1) ^ 1);  // previous occupant of the slot fully drained
        tc_fence_after();
        const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + qslot * ATT_QBYTES), 16, 1024);
        for (int h = 0; h < p.parts; ++h) {
          const uint32_t t_s = tmem_base + (slot * p.parts + h) * p.region;
          const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + h * p.half * ATT_D * 2), 16, 1024);
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k) mma_f16_ss(t_s, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0);
          if (h == p.parts - 1) tc_commit(&q_empty[qslot]);  // Q tile consumed once these MMAs retire
          tc_commit(&s_full[slot * p.parts + h]);
        }
        if (last_of_item) tc_commit(k_empty);  // K may be replaced once the last S product of the item has retired
      };
      auto issue_pv = [&](int n, bool first_of_item, bool last_of_item) {
        const int slot = p.nbuf == 2 ? (n & 1) : 0;
        const uint32_t use = p.nbuf == 2 ? (uint32_t)(n >> 1) : (uint32_t)n;
        if (first_of_item) mbar_wait(v_full, kv_ph);
        for (int h = 0; h < p.parts; ++h) {
          mbar_wait(&p_full[slot * p.parts + h], use & 1);
          tc_fence_after();
          const uint32_t t_s = tmem_base + (slot * p.parts + h) * p.region;
          // V is [key][d] = MN-major B operand: 8-key groups are 1024 B apart (SBO); one 64-wide MN atom.
          const uint64_t v_desc = make_sdesc_sw128(smem_u32(sV + h * p.half * ATT_D * 2), 1024, 1024);
          for (int k = 0; k < p.half / 16; ++k) {
            // 16 keys per MMA = 8 TMEM columns of P and 2048 B (= 128 in >>4 units) of V
            mma_f16_ts(t_s + p.o_col, t_s + 8 * k, v_desc + 128 * k, idesc_o, k != 0);
          }
          tc_commit(&o_full[slot * p.parts + h]);
        }
        if (last_of_item) tc_commit(v_empty);  // V may be replaced once every MMA of the item has retired
      };
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
        mbar_wait(k_full, kv_ph);
        tc_fence_after();
        if (p.nbuf == 2) {
          // S of tile n+1 is issued before P V of tile n: the tensor pipe runs ahead of the softmax warps
          issue_s(qn, p.tq == 1);
          for (int t = 0; t < p.tq; ++t) {
            if (t + 1 < p.tq) issue_s(qn + t + 1, t + 2 == p.tq);
            issue_pv(qn + t, t == 0, t + 1 == p.tq);
          }
        } else {
          for (int t = 0; t < p.tq; ++t) {
            issue_s(qn + t, t + 1 == p.tq);
            issue_pv(qn + t, t == 0, t + 1 == p.tq);
          }
        }
        kv_ph ^= 1;
        qn += p.tq;
      }
    }
  } else {
    // ===================== softmax + epilogue (warps 2..9) =====================
    const int q = warp & 3;              // TMEM lane quarter
    const int g = (warp - 2) >> 2;       // key half handled by this warp group
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int half = p.half;
    int qn = 0;
    for (int item = blockIdx.x; item < p.num_items; item += gridDim.x) {
      const int fh = item / p.items_per_head;
      const int frame = fh / p.heads, head = fh - frame * p.heads;
      const int qt0 = (item - fh * p.items_per_head) * p.tq;
      for (int t = 0; t < p.tq; ++t, ++qn) {
        const int slot = p.nbuf == 2 ? (qn & 1) : 0;
        const uint32_t use = p.nbuf == 2 ? (uint32_t)(qn >> 1) : (uint32_t)qn;
        const int passes = p.parts >> 1;
        ATT_STAMP(0);
#pragma unroll 1
        for (int pass = 0; pass < passes; ++pass) {
        const int part = 2 * pass + g;
        const uint32_t t_mine = tmem_base + (slot * p.parts + part) * p.region + lane_off;
        mbar_wait(&s_full[slot * p.parts + part], use & 1);
        tc_fence_after();
        if (pass == 0) ATT_STAMP(1);
        // ---- two sweeps over this part of S: exact row maximum, then probabilities ----
        // TMEM -> register reads are cheap (tools/micro/tmem_rate.cu: ~960 B/clk per SM with 8 warps, i.e. a 128 x 512
        // fp32 score tile in ~270 clocks — round 1 assumed 64 B/clk and built a single-sweep, lazily-raised shift around
        // that: an fp16-overflow vote after every 32 columns, a rescale path and a redo path, all inlined twice per pass.
        // The ncu source page of that form showed the softmax warps stalled on instruction fetch (stall_no_inst: a
        // 7 000-instruction kernel) and on the per-chunk vote chain, not on MUFU or TMEM). With the exact maximum first,
        // every probability is <= 1: no overflow, no vote, no rescale, and the hot loop is ~130 instructions.
        float m_scaled;
        float sum;
        {
          uint32_t va[32], vb[32];
          auto max32 = [&](const uint32_t (&v)[32]) {
            float c0 = __uint_as_float(v[0]), c1 = __uint_as_float(v[1]), c2 = __uint_as_float(v[2]),
                  c3 = __uint_as_float(v[3]);
#pragma unroll
            for (int j = 4; j < 32; j += 4) {
              c0 = fmaxf(c0, __uint_as_float(v[j])); c1 = fmaxf(c1, __uint_as_float(v[j + 1]));
              c2 = fmaxf(c2, __uint_as_float(v[j + 2])); c3 = fmaxf(c3, __uint_as_float(v[j + 3]));
            }
            return fmaxf(fmaxf(c0, c1), fmaxf(c2, c3));
          };
          // sweep 1: exact maximum (scale > 0, so max(s) * c = max(s * c))
          float mx = -INFINITY;
          tmem_ld32(t_mine, va);
#pragma unroll 1
          for (int c = 0; c < half; c += 64) {
            tmem_ld_wait();
            if (c + 32 < half) tmem_ld32(t_mine + c + 32, vb);
            mx = fmaxf(mx, max32(va));
            if (c + 32 < half) {
              tmem_ld_wait();
              if (c + 64 < half) tmem_ld32(t_mine + c + 64, va);
              mx = fmaxf(mx, max32(vb));
            }
          }
          m_scaled = mx * p.scale_log2;
          // sweep 2: p = 2^(s c - m) <= 1 as packed fp16 pairs (= the MMA operand of P V), written back over S columns
          // that have been consumed; row sum through a 16-leaf fp16 tree per 32 columns, widened to fp32.
          // Per pair of scores: two FFMA, two fp32 MUFU exp2, one cvt.rn.f16x2 pack (24 elements / clk / SM; the packed
          // ex2.approx.f16x2 round 1 used runs at a quarter of the fp32 instruction's rate: tools/micro/exp_rate.cu).
          float s0 = 0.f;
          auto emit = [&](const uint32_t (&v)[32], int c) {
            uint32_t ps[16];
#pragma unroll
            for (int j = 0; j < 16; ++j)
              ps[j] = pack_f16x2(ex2_f32(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -m_scaled)),
                                 ex2_f32(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -m_scaled)));
            uint32_t t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = add_f16x2(ps[2 * j], ps[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j) t[j] = add_f16x2(t[2 * j], t[2 * j + 1]);
            const float2 a = unpack_f16x2(add_f16x2(t[0], t[1])), b = unpack_f16x2(add_f16x2(t[2], t[3]));
            s0 += (a.x + a.y) + (b.x + b.y);
            tmem_st16(t_mine + (c >> 1), ps);  // columns [c/2, c/2+16): below every column still to be read
          };
          tmem_ld32(t_mine, va);
#pragma unroll 1
          for (int c = 0; c < half; c += 64) {
            tmem_ld_wait();
            if (c + 32 < half) tmem_ld32(t_mine + c + 32, vb);
            emit(va, c);
            if (c + 32 < half) {
              tmem_ld_wait();
              if (c + 64 < half) tmem_ld32(t_mine + c + 64, va);
              emit(vb, c + 32);
            }
          }
          sum = s0;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[slot * p.parts + part]);
        // (shift, sum) of this key part -> shared memory. The previous tile's readers of this slot are past their
        // named barrier 3, which every softmax warp crosses after its last read.
        s_stat[slot * (2 * 2 * ATT_BM) + (part * 2 + 0) * ATT_BM + row_in_tile] = m_scaled;
        s_stat[slot * (2 * 2 * ATT_BM) + (part * 2 + 1) * ATT_BM + row_in_tile] = sum;
        }  // pass
        ATT_STAMP(2);

        // ---- exchange (max, sum) of every key part ----
        float* st = s_stat + slot * (2 * 2 * ATT_BM);   // [part][m, l][128]; parts = 4 only occurs with one slot
        // (the thread that issues the output TMA stores first makes sure the previous tile's store has read the staging
        // tile: everybody may overwrite it after the barrier)
        if (threadIdx.x == 64) bulk_wait_group_read<0>();
        asm volatile("bar.sync 2, 256;" ::: "memory");
        float m_all = -INFINITY;
#pragma unroll
        for (int pt = 0; pt < 4; ++pt)
          if (pt < p.parts) m_all = fmaxf(m_all, st[(pt * 2 + 0) * ATT_BM + row_in_tile]);
        float wgt[4] = {0.f, 0.f, 0.f, 0.f};
        float l_all = 0.f;
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          if (pt < p.parts) {
            wgt[pt] = ex2_approx(st[(pt * 2 + 0) * ATT_BM + row_in_tile] - m_all);
            l_all = fmaf(wgt[pt], st[(pt * 2 + 1) * ATT_BM + row_in_tile], l_all);
          }
        }
        const float inv = 1.0f / l_all;
        if (p.stats != nullptr && g == 0) {
          const int qrow_s = (qt0 + t) * ATT_BM + row_in_tile;
          if (qrow_s < p.Sq) {
            float* sp = p.stats + ((static_cast<int64_t>(frame) * p.Sq + qrow_s) * p.heads + head) * 2;
            sp[0] = m_all;
            sp[1] = l_all;
          }
        }

        // ---- O = sum_part a_part O_part : this group normalises and stores output columns [32 g, 32 g + 32) ----
        ATT_STAMP(3);
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
        for (int pt = 0; pt < 4; ++pt) {
          if (pt >= p.parts) break;
          mbar_wait(&o_full[slot * p.parts + pt], use & 1);
          tc_fence_after();
          uint32_t ov[32];
          tmem_ld32(tmem_base + (slot * p.parts + pt) * p.region + lane_off + p.o_col + 32 * g, ov);
          tmem_ld_wait();
          const float a = wgt[pt] * inv;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = fmaf(a, __uint_as_float(ov[j]), acc[j]);
        }
        ATT_STAMP(4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[slot]);  // slot free for the next S as soon as O is in registers
        // O tile -> swizzled staging -> ONE asynchronous TMA store per tile (direct global stores kept these warps, which
        // also are the softmax warps of the next tile, busy for ~0.9 us per tile). Group g owns bytes [64 g, 64 g + 64)
        // of every 128-byte row = 16-byte chunks 4 g .. 4 g + 3.
        {
          const uint32_t srow = smem_u32(sO) + (uint32_t)row_in_tile * 128u;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float* r = acc + 8 * j;
            st_shared_v4(srow + ((uint32_t)((4 * g + j) ^ (row_in_tile & 7)) << 4), pack_bf16x2(r[0], r[1]),
                         pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 3, 256;" ::: "memory");
          if (threadIdx.x == 64) {
            tma_store_2d(&tmO, sO, head * ATT_D, frame * p.Sq + (qt0 + t) * ATT_BM);
            bulk_commit_group();
          }
        }
        ATT_STAMP(5);
      }
    }
  }

  if (threadIdx.x == 64) bulk_wait_group<0>();   // the staging tile must outlive the last store
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_d64(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
             int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, cudaStream_t stream) {
  return attn_d64_chunk(Q, ldq, K, ldk, V, ldv, O, ldo, frames, heads, Sq, Skv, Skv, nullptr, scale, stream);
}

int attn_d64_chunk(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                   int64_t ldo, int frames, int heads, int Sq, int Skv, int kv_frame_rows, float* stats, float scale,
                   cudaStream_t stream) {
  RALD_REQUIRE(frames > 0 && heads > 0 && Sq > 0, "attn: bad sizes");
  if (Skv == 512 && g_attn_dbg == nullptr && attn_streams_enabled())
    return attn_d64_streams(Q, ldq, K, ldk, V, ldv, O, ldo, frames, heads, Sq, kv_frame_rows, stats, scale, stream);
  RALD_REQUIRE(kv_frame_rows >= Skv, "attn: %d key rows per frame < Skv=%d", kv_frame_rows, Skv);
  RALD_REQUIRE(Skv >= 64 && Skv <= 512 && Skv % 64 == 0, "attn: Skv=%d must be a multiple of 64 in [64, 512]", Skv);
  RALD_REQUIRE(Sq % ATT_BM == 0, "attn: Sq=%d must be a multiple of 128", Sq);
  RALD_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "attn: output not 16-byte aligned");
  RALD_REQUIRE((uint64_t)ldo * 2 % 16 == 0, "attn: output row pitch must be a multiple of 16 bytes");
  CUtensorMap tmQ, tmK, tmV, tmO;
  RALD_TRY(make_tmap_out(&tmO, O, (uint64_t)frames * Sq, (uint64_t)heads * ATT_D, (uint64_t)ldo, false, ATT_BM));
  const int parts = Skv == 512 ? 4 : 2;
  const uint32_t kv_box = Skv / parts;  // one box per key part (<= 256 rows)
  RALD_TRY(make_tmap_2d_bf16(&tmQ, Q, (uint64_t)frames * Sq, (uint64_t)heads * ATT_D, (uint64_t)ldq, ATT_BM));
  // (K / V may point at a key chunk inside each frame's rows: the last frame's chunk ends Skv rows after its start)
  const uint64_t kv_rows = (uint64_t)(frames - 1) * kv_frame_rows + Skv;
  RALD_TRY(make_tmap_2d_bf16(&tmK, K, kv_rows, (uint64_t)heads * ATT_D, (uint64_t)ldk, kv_box));
  RALD_TRY(make_tmap_2d_bf16(&tmV, V, kv_rows, (uint64_t)heads * ATT_D, (uint64_t)ldv, kv_box));
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.Sq = Sq;
  p.Skv = Skv;
  p.frames = frames;
  p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.kv_frame_rows = kv_frame_rows;
  p.stats = stats;
  p.parts = parts;
  p.half = Skv / parts;
  // part region: S occupies [0, half), P (fp16 pairs) [0, half/2), O (64 fp32 columns) right after P
  p.o_col = p.half / 2 < 32 ? 32u : (uint32_t)(p.half / 2);
  uint32_t need = p.o_col + ATT_D;
  if (need < (uint32_t)p.half) need = p.half;
  uint32_t region = 32;
  while (region < need) region <<= 1;
  p.region = region;
  p.nbuf = 2 * parts * region <= 512 ? 2 : 1;
  p.q_tiles = Sq / ATT_BM;
  // query tiles per work item: rounds x (K/V load + tq tiles), in units of one tile
  const int sms = device_sm_count();
  const double kv_cost = 0.5 * Skv / 512.0;
  int best_tq = 1;
  double best = -1.0;
  for (int tq = 1; tq <= p.q_tiles; ++tq) {
    if (p.q_tiles % tq != 0) continue;
    const long items = (long)frames * heads * (p.q_tiles / tq);
    const double cost = (double)((items + sms - 1) / sms) * (kv_cost + tq);
    if (best < 0 || cost < best - 1e-9) { best = cost; best_tq = tq; }
  }
  p.tq = best_tq;
  p.items_per_head = p.q_tiles / p.tq;
  p.num_items = frames * heads * p.items_per_head;
  p.dbg = g_attn_dbg;
  const int smem_bytes = 3 * ATT_QBYTES + 2 * Skv * ATT_D * 2 + 2 * 2 * 2 * ATT_BM * 4 + 1024 + 256;
  static int configured_bytes = 0;
  if (smem_bytes > configured_bytes) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(attn_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured_bytes = smem_bytes;
  }
  const int grid = p.num_items < sms ? p.num_items : sms;
  ProfScope prof(FAM_ATTN, stream, 4.0 * frames * heads * Sq * Skv * ATT_D);
  RALD_CHECK_CUDA(launch_pdl(attn_d64_kernel, dim3(grid), dim3(ATT_THREADS), smem_bytes, stream, tmQ, tmK, tmV, tmO, p));
  RALD_LAUNCHED();
  return 0;
}

int attn_merge_chunks(const void* o_chunks, int64_t chunk_stride, int64_t ldc, const float* stats,
                      int64_t stats_chunk_stride, int chunks, int heads, int64_t rows, void* out, int64_t ldo,
                      cudaStream_t stream);

int attn_d64_long(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                  int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, void* o_chunks, float* stats,
                  cudaStream_t stream) {
  RALD_REQUIRE(Skv % 512 == 0 && Skv >= 1024 && Skv <= 4096, "attn_long: Skv=%d must be a multiple of 512 in [1024, 4096]",
               Skv);
  RALD_REQUIRE(o_chunks != nullptr && stats != nullptr, "attn_long: scratch buffers missing");
  const int chunks = Skv / 512;
  const int64_t rows = (int64_t)frames * Sq;
  const int64_t wdt = (int64_t)heads * ATT_D;
  const __nv_bfloat16* Kb = reinterpret_cast<const __nv_bfloat16*>(K);
  const __nv_bfloat16* Vb = reinterpret_cast<const __nv_bfloat16*>(V);
  __nv_bfloat16* oc = reinterpret_cast<__nv_bfloat16*>(o_chunks);
  for (int c = 0; c < chunks; ++c)
    RALD_TRY(attn_d64_chunk(Q, ldq, Kb + (int64_t)c * 512 * ldk, ldk, Vb + (int64_t)c * 512 * ldv, ldv, oc + c * rows * wdt,
                            wdt, frames, heads, Sq, 512, Skv, stats + c * rows * heads * 2, scale, stream));
  return attn_merge_chunks(oc, rows * wdt, wdt, stats, rows * heads * 2, chunks, heads, rows, O, ldo, stream);
}

// Long contexts (Skv > 512): the keys are processed in chunks of <= 512 by attn_d64_chunk, each chunk yielding its
// own normalised output O_c (bf16) and statistics (m_c, l_c); the exact softmax over all keys is
//   O = sum_c w_c O_c / sum_c w_c,   w_c = l_c 2^(m_c - max_c m_c).
// One thread per (row, head, 8 columns).
__global__ void __launch_bounds__(256)
attn_merge_kernel(const __nv_bfloat16* __restrict__ o_chunks, int64_t chunk_stride, int64_t ldc,
                  const float* __restrict__ stats, int64_t stats_chunk_stride, int chunks, int heads, int64_t rows,
                  __nv_bfloat16* __restrict__ out, int64_t ldo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row, head, octet)
  const int64_t total = rows * heads * 8;
  if (i >= total) return;
  const int oct = (int)(i & 7);
  const int head = (int)((i >> 3) % heads);
  const int64_t row = i / (8 * heads);
  float m[8], l[8];
  float mx = -INFINITY;
  for (int c = 0; c < chunks; ++c) {
    const float* sp = stats + c * stats_chunk_stride + (row * heads + head) * 2;
    m[c] = sp[0];
    l[c] = sp[1];
    mx = fmaxf(mx, m[c]);
  }
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float wsum = 0.f;
  for (int c = 0; c < chunks; ++c) {
    const float w = l[c] * ex2_approx(m[c] - mx);
    wsum += w;
    const uint4 v = *reinterpret_cast<const uint4*>(o_chunks + c * chunk_stride + row * ldc + head * ATT_D + oct * 8);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] = fmaf(w, __uint_as_float(u[j] << 16), acc[2 * j]);
      acc[2 * j + 1] = fmaf(w, __uint_as_float(u[j] & 0xffff0000u), acc[2 * j + 1]);
    }
  }
  const float inv = 1.0f / wsum;
  *reinterpret_cast<uint4*>(out + row * ldo + head * ATT_D + oct * 8) =
      make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv),
                 pack_bf16x2(acc[4] * inv, acc[5] * inv), pack_bf16x2(acc[6] * inv, acc[7] * inv));
}

int attn_merge_chunks(const void* o_chunks, int64_t chunk_stride, int64_t ldc, const float* stats,
                      int64_t stats_chunk_stride, int chunks, int heads, int64_t rows, void* out, int64_t ldo,
                      cudaStream_t stream) {
  RALD_REQUIRE(chunks >= 1 && chunks <= 8, "attn_merge: %d chunks (1..8)", chunks);
  RALD_REQUIRE(ldc % 8 == 0 && ldo % 8 == 0 && chunk_stride % 8 == 0, "attn_merge: strides must be 16-byte multiples");
  const int64_t total = rows * heads * 8;
  ProfScope prof(FAM_ATTN, stream, 0.0);
  attn_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(o_chunks), chunk_stride, ldc, stats, stats_chunk_stride, chunks, heads, rows,
      reinterpret_cast<__nv_bfloat16*>(out), ldo);
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

// softmax(Q K^T * scale) V over Skv = chunks * 512 keys per frame (K / V rows frame * Skv + key): the chunked form of
// rald_attn_d64 for contexts longer than TMEM holds. o_chunks: bf16 scratch [chunks][frames*Sq][heads*64];
// stats: fp32 scratch [chunks][frames*Sq][heads][2].
extern "C" int rald_attn_d64_long(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv,
                                  void* O, int64_t ldo, int frames, int heads, int Sq, int Skv, float scale,
                                  void* o_chunks, float* stats, void* stream) {
  return rald::attn_d64_long(Q, ldq, K, ldk, V, ldv, O, ldo, frames, heads, Sq, Skv, scale, o_chunks, stats,
                             static_cast<cudaStream_t>(stream));
}

extern "C" int rald_attn_debug_buffer(unsigned long long* dev_buf) {
  rald::g_attn_dbg = dev_buf;
  return 0;
}
