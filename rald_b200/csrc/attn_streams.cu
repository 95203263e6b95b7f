// softmax(Q K^T * scale) V for head_dim 64 and Skv = 512 (the denoiser's self-attention, the 512-key chunks of longer
// contexts, the training forward): FOUR INDEPENDENT SOFTMAX STREAMS per CTA instead of one tile at a time.
//
// attn.cu holds the whole 128 x 512 score tile of ONE query tile in TMEM (all 512 columns), so the tensor pipe idles
// while the softmax warps work and the softmax warps idle while the P V products drain: 6.3 us per tile, of which 5.4 us
// are the exponentials of 8 warps that issue in 16 % of the cycles (profiles/r02_attn_stalls.md). Here TMEM is cut into
// four stream regions of 128 columns = [S / P chunk: 64 columns | O accumulator: 64 columns], each owned by four warps
// (thread <-> query row <-> TMEM lane). A stream walks over 64-key chunks:
//     S = Q K_c^T (MMA)  ->  row max, lazily raised shift, p = 2^(s c - m) as fp16 pairs over S  ->  O += P V_c (MMA)
// and while one stream waits for its MMAs the other three keep the exp2 unit busy; the single MMA thread serves the
// streams round-robin. Row statistics are thread-local (no shuffles, no shared-memory exchange inside a tile).
// A query tile is always served by TWO streams (even / odd chunks), so its arithmetic does not depend on the batch:
//   tq = 2: two query tiles of a (frame, head) in flight = four streams     [batches that fill the GPU]
//   tq = 1: one query tile, two streams; the other two stay idle             [fewer tiles than SMs]
// The streams of a tile keep their own (shift, sum, O) and are merged exactly at the end of the tile (as attn.cu merges
// its key parts). The shift of a stream is the running maximum raised lazily: it moves only when a chunk maximum exceeds
// it by more than 2^8 (p <= 256 in fp16), which rescales that stream's O in TMEM (rare; exact either way).
//   warp 0 lane 0 : TMA producer (Q tile ring of 4, K / V as eight 64-key chunks each, recycled chunk by chunk so the
//                   next item's keys arrive while the current item finishes)
//   warp 1 lane 0 : tcgen05 issuer
//   warps 2..17   : stream s = (warp - 2) / 4, TMEM lane quarter = warp % 4
// The output tile is staged (bf16, 128-byte swizzled rows) in the tile's own Q slot — dead once the last S product of
// the tile has retired — and leaves through one TMA store. Reference: model/models_radar_generation.py:66-75.
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int AS_BM = 128;                     // query rows per tile
constexpr int AS_D = 64;                       // head dim
constexpr int AS_CK = 64;                      // keys per chunk
constexpr int AS_SKV = 512;
constexpr int AS_NCH = AS_SKV / AS_CK;         // 8 chunks
constexpr int AS_STREAMS = 4;
constexpr int AS_THREADS = 64 + AS_STREAMS * 128 + 32;   // TMA warp, MMA warp A, 16 stream warps, MMA warp B
constexpr int AS_QBYTES = AS_BM * AS_D * 2;    // 16 KB
constexpr int AS_CBYTES = AS_CK * AS_D * 2;    // 8 KB: one K or V chunk
constexpr int AS_QSLOTS = 4;
constexpr float AS_RAISE = 8.0f;               // log2 units

struct AttnStreamParams {
  int Sq, frames, heads;
  int tq;              // query tiles per work item (1 or 2)
  int items_per_head;  // (Sq / 128) / tq
  int num_items;
  float scale_log2;    // scale * log2(e)
  int kv_frame_rows;   // rows per frame in the K / V tensors (>= 512; a key CHUNK of a longer context when larger)
  int skew;            // tq = 2: the second tile slot starts behind the first
  float* stats;        // optional [frames*Sq][heads][2] = (shift, sum) of the keys seen by this call
  unsigned long long* dbg;   // optional %globaltimer stamps of CTA 0 (tools/attn_streams_phases.py): issuers [n < 32][stream]
                             // [PV issued, S issued], streams 256 + [stream][n < 32][S seen, max done, P written]
};
static unsigned long long* g_attn_streams_dbg = nullptr;
#define AS_STAMP_I(n_, s_, e_)                                                                      \
  do {                                                                                              \
    if (p.dbg != nullptr && blockIdx.x == 0 && (n_) < 32) p.dbg[((n_) * 4 + (s_)) * 2 + (e_)] = global_timer_ns(); \
  } while (0)
#define AS_STAMP_S(n_, e_)                                                                          \
  do {                                                                                              \
    if (p.dbg != nullptr && blockIdx.x == 0 && (n_) < 32 && q == 0 && lane == 0)                    \
      p.dbg[256 + (s * 32 + (n_)) * 3 + (e_)] = global_timer_ns();                                  \
  } while (0)

template <int TQ>
__global__ void __launch_bounds__(AS_THREADS, 1)
attn_d64_streams_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                        const AttnStreamParams p) {
  // Two streams per tile whatever TQ: a tile's arithmetic (even / odd chunks, merge order) does not depend on the batch,
  // so a frame computes bit-identical values alone and inside a batch. TQ = 1 leaves streams 2, 3 and MMA warp B idle.
  constexpr int NS = 2;                   // streams per tile
  constexpr int ACTIVE = NS * TQ;         // streams in use
  constexpr int NCH = AS_NCH / NS;        // chunks per stream and tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // [4][128 x 64] (slot = tile counter & 3; later the O staging tile)
  uint8_t* sK = sQ + AS_QSLOTS * AS_QBYTES;             // [8][64 x 64]
  uint8_t* sV = sK + AS_NCH * AS_CBYTES;                // [8][64 x 64]
  float* s_stat = reinterpret_cast<float*>(sV + AS_NCH * AS_CBYTES);   // [stream 4][shift, sum, max][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + AS_STREAMS * 3 * AS_BM);
  uint64_t* q_full = bars + 0;      // [4]
  uint64_t* q_empty = bars + 4;     // [4]
  uint64_t* kv_full = bars + 8;     // [8]
  uint64_t* kv_empty = bars + 16;   // [8]
  uint64_t* s_full = bars + 24;     // [4]
  uint64_t* p_full = bars + 28;     // [4]
  uint64_t* o_full = bars + 32;     // [4]
  uint64_t* o_free = bars + 36;     // [2] per tile slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    for (int i = 0; i < AS_QSLOTS; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < AS_NCH; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], TQ);   // tq = 2: one commit per MMA warp (each serves one tile slot)
    }
    for (int i = 0; i < AS_STREAMS; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&o_free[i], 4 * NS);
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
      uint32_t it = 0;   // items of this CTA
      uint32_t qn = 0;   // query tiles of this CTA
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
        const int fh = item / p.items_per_head;
        const int frame = fh / p.heads, head = fh - frame * p.heads;
        const int qt0 = (item - fh * p.items_per_head) * TQ;
        const int kv_row0 = frame * p.kv_frame_rows;
        for (int t = 0; t < TQ; ++t, ++qn) {
          const int slot = qn & 3;
          mbar_wait(&q_empty[slot], ((qn >> 2) & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[slot], AS_QBYTES);
          tma_load_2d(sQ + slot * AS_QBYTES, &tmQ, &q_full[slot], head * AS_D, frame * p.Sq + (qt0 + t) * AS_BM);
        }
        for (int c = 0; c < AS_NCH; ++c) {
          mbar_wait(&kv_empty[c], (it & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[c], 2 * AS_CBYTES);
          tma_load_2d(sK + c * AS_CBYTES, &tmK, &kv_full[c], head * AS_D, kv_row0 + c * AS_CK);
          tma_load_2d(sV + c * AS_CBYTES, &tmV, &kv_full[c], head * AS_D, kv_row0 + c * AS_CK);
        }
      }
    }
  } else if (warp == 1 || warp == 2 + 4 * AS_STREAMS) {
    // ===================== MMA issuers: warp 1 serves streams 0, 1; the last warp streams 2, 3 =====================
    // (one issuer for all four streams was the pace: an mbarrier probe + four MMAs cost it ~0.3 us, 0.62 us per stream
    // step, 2.5 us per round of the four streams whose softmax takes 0.55 us: tools/attn_streams_phases.py)
    // The whole warp runs this code with warp-uniform values and ONE ELECTED lane issues: with the loop under `lane == 0`
    // the compiler cannot keep descriptors and addresses in uniform registers and wraps every tcgen05.mma / commit in an
    // ELECT / R2UR.BROADCAST loop of ~14 instructions (~120 clocks per 48-clock MMA: 0.6 us per issuer step whatever
    // the number of MMAs or mbarrier probes; tools/attn_streams_phases.py, cuobjdump -sass).
    if (TQ == 2 || warp == 1) {
      const int s_lo = __shfl_sync(0xffffffffu, warp == 1 ? 0 : 2, 0);
      const uint32_t tmem_base_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t idesc_s = make_idesc(FMT_BF16, AS_BM, AS_CK, 0, 0);   // Q (smem) x K_c (smem) -> 128 x 64 fp32
      const uint32_t idesc_o = make_idesc(FMT_F16, AS_BM, AS_D, 0, 1);     // P fp16 (TMEM) x V_c fp16 (smem, MN-major)
      uint32_t it = 0;
      uint32_t n = 0;    // chunks issued per stream so far (the streams advance in lock step)
      // the second tile slot starts half a chunk cycle behind the first: in lock step all four streams are in their
      // exponentials (XU-bound) and then all wait for their MMAs; shifted, one slot's exponentials run under the other's MMAs
      if (TQ == 2 && s_lo != 0 && p.skew != 0 && (int)(blockIdx.x + gridDim.x) < p.num_items) mbar_wait(&p_full[0], 0);
      for (int item = blockIdx.x; item < p.num_items; item += gridDim.x, ++it) {
#pragma unroll 1
        for (int k = 0; k <= NCH; ++k) {
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int s = s_lo + s2;
            const int ts = s / NS, sub = s % NS;
            const uint32_t t_sp = tmem_base_u + s * 128;   // S / P chunk
            const uint32_t t_o = t_sp + 64;              // O accumulator
            if (k >= 1) {
              // ---- O_s (+)= P V_c for the chunk whose probabilities the stream has just written ----
              const int c = (k - 1) * NS + sub;
              mbar_wait(&p_full[s], (n + k - 1) & 1);
              if (k == 1) mbar_wait(&o_free[ts], (it & 1) ^ 1);   // the previous tile's O has been read
              tc_fence_after();
              // V chunk is [key][d] = MN-major B operand: 8-key groups are 1024 B apart (SBO); one 64-wide MN atom
              const uint64_t v_desc = make_sdesc_sw128(smem_u32(sV + c * AS_CBYTES), 1024, 1024);
              if (elect_one()) {
#pragma unroll
                for (int j = 0; j < AS_CK / 16; ++j)   // 16 keys per MMA = 8 TMEM columns of P and 2048 B of V
                  mma_f16_ts(t_o, t_sp + 8 * j, v_desc + 128 * j, idesc_o, k > 1 || j > 0);
                if (k == NCH) tc_commit(&o_full[s]);
                AS_STAMP_I(n + k - 1, s, 0);
                tc_commit(&kv_empty[c]);
              }
              __syncwarp();
            }
            // ---- S_s = Q K_c^T of the stream's next chunk (in order behind the P V product that read P). The first chunk
            // of the NEXT item follows the last P V product of this one: it runs under the tile's merge / store ----
            const bool next_item = (k == NCH) && (item + (int)gridDim.x < p.num_items);
            if ((k < NCH && (k > 0 || it == 0)) || next_item) {
              const uint32_t it_s = next_item ? it + 1 : it;
              const int c = (next_item ? 0 : k) * NS + sub;
              const uint32_t qn = it_s * TQ + ts;
              const int qslot = qn & 3;
              if ((k == 0 || next_item) && (sub & 1) == 0) mbar_wait(&q_full[qslot], (qn >> 2) & 1);   // once per issuer and tile
              mbar_wait(&kv_full[c], it_s & 1);
              tc_fence_after();
              const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + qslot * AS_QBYTES), 16, 1024);
              const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + c * AS_CBYTES), 16, 1024);
              if (elect_one()) {
#pragma unroll
                for (int j = 0; j < AS_D / 16; ++j) mma_f16_ss(t_sp, q_desc + 2 * j, k_desc + 2 * j, idesc_s, j != 0);
                tc_commit(&s_full[s]);
                AS_STAMP_I(n + (next_item ? NCH : k), s, 1);
              }
              __syncwarp();
            }
          }
        }
        n += NCH;
      }
    }
  } else if (warp < 2 + 4 * AS_STREAMS) {
    // ===================== softmax streams =====================
    const int s = (warp - 2) >> 2;
    const int ts = s / NS, sub = s % NS;
    const bool active = s < ACTIVE;
    const int q = warp & 3;                  // TMEM lane quarter
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t t_sp = tmem_base + s * 128 + lane_off;
    const uint32_t t_o = t_sp + 64;
    const bool storer = (threadIdx.x == 64 + ts * NS * 128);   // issues this tile slot's TMA stores
    int pending_slot = -1;                   // Q slot whose output store may still be reading shared memory
    uint32_t it = 0, n = 0;
    for (int item = blockIdx.x; active && item < p.num_items; item += gridDim.x, ++it) {
      const int fh = item / p.items_per_head;
      const int frame = fh / p.heads, head = fh - frame * p.heads;
      const int tile = (item - fh * p.items_per_head) * TQ + ts;
      const int qslot = (it * TQ + ts) & 3;
      float m_used = -INFINITY, m_true = -INFINITY, l = 0.f;   // shift in use, exact running maximum, sum
#pragma unroll 1
      for (int k = 0; k < NCH; ++k, ++n) {
        mbar_wait(&s_full[s], n & 1);
        tc_fence_after();
        AS_STAMP_S(n, 0);
        uint32_t va[32], vb[32];
        // ---- sweep 1: chunk maximum ----
        tmem_ld32(t_sp, va);
        tmem_ld32(t_sp + 32, vb);
        tmem_ld_wait();
        float mc;
        {
          float c0 = __uint_as_float(va[0]), c1 = __uint_as_float(va[1]), c2 = __uint_as_float(va[2]),
                c3 = __uint_as_float(va[3]);
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            c0 = fmaxf(c0, __uint_as_float(va[j])); c1 = fmaxf(c1, __uint_as_float(va[j + 1]));
            c2 = fmaxf(c2, __uint_as_float(va[j + 2])); c3 = fmaxf(c3, __uint_as_float(va[j + 3]));
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            c0 = fmaxf(c0, __uint_as_float(vb[j])); c1 = fmaxf(c1, __uint_as_float(vb[j + 1]));
            c2 = fmaxf(c2, __uint_as_float(vb[j + 2])); c3 = fmaxf(c3, __uint_as_float(vb[j + 3]));
          }
          mc = fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)) * p.scale_log2;   // scale > 0
        }
        m_true = fmaxf(m_true, mc);
        AS_STAMP_S(n, 1);
        if (k == 0) {
          m_used = mc;
        } else {
          const bool raise = mc > m_used + AS_RAISE;
          if (__any_sync(0xffffffffu, raise)) {
            // the P V product of the previous chunk has retired (s_full is committed behind it): O_s is at rest
            const float f = raise ? ex2_f32(m_used - mc) : 1.0f;
            if (raise) m_used = mc;
            l *= f;
#pragma unroll 1
            for (int h = 0; h < AS_D / 16; ++h) {
              uint32_t o[16];
              tmem_ld16(t_o + 16 * h, o);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
              tmem_st16(t_o + 16 * h, o);
            }
          }
        }
        // ---- sweep 2: p = 2^(s c - m) as packed fp16 pairs over the consumed S columns; row sum in fp32 ----
        const float msh = m_used;
        float s0 = 0.f;
        auto emit = [&](const uint32_t (&v)[32], int col) {
          uint32_t ps[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            ps[j] = pack_f16x2(ex2_f32(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -msh)),
                               ex2_f32(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -msh)));
          uint32_t t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] = add_f16x2(ps[2 * j], ps[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = add_f16x2(t[2 * j], t[2 * j + 1]);
          const float2 a = unpack_f16x2(add_f16x2(t[0], t[1])), b = unpack_f16x2(add_f16x2(t[2], t[3]));
          s0 += (a.x + a.y) + (b.x + b.y);
          tmem_st16(t_sp + (col >> 1), ps);
        };
        emit(va, 0);     // both halves are in registers: P may overwrite any S column
        emit(vb, 32);
        l += s0;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[s]);
        AS_STAMP_S(n, 2);
        // the output store of this tile slot's previous tile has long finished reading its staging tile: release the Q slot
        if (k == 0 && storer && pending_slot >= 0) {
          bulk_wait_group_read<0>();
          mbar_arrive(&q_empty[pending_slot]);
          pending_slot = -1;
        }
      }

      // ---- merge the NS streams of this tile: exchange (shift, sum) ----
      s_stat[(s * 3 + 0) * AS_BM + row_in_tile] = m_used;
      s_stat[(s * 3 + 1) * AS_BM + row_in_tile] = l;
      s_stat[(s * 3 + 2) * AS_BM + row_in_tile] = m_true;
      asm volatile("bar.sync %0, %1;" ::"r"(2 + ts), "n"(NS * 128) : "memory");
      float m_all = -INFINITY;
#pragma unroll
      for (int j = 0; j < NS; ++j) m_all = fmaxf(m_all, s_stat[((ts * NS + j) * 3 + 0) * AS_BM + row_in_tile]);
      float wgt[NS];
      float l_all = 0.f;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        wgt[j] = ex2_f32(s_stat[((ts * NS + j) * 3 + 0) * AS_BM + row_in_tile] - m_all);
        l_all = fmaf(wgt[j], s_stat[((ts * NS + j) * 3 + 1) * AS_BM + row_in_tile], l_all);
      }
      const float inv = 1.0f / l_all;
      if (p.stats != nullptr && sub == 0) {
        const int qrow = tile * AS_BM + row_in_tile;
        float* sp = p.stats + ((static_cast<int64_t>(frame) * p.Sq + qrow) * p.heads + head) * 2;
        // reported against the EXACT row maximum (as attn.cu does): sum_j 2^(s_j c - max)
        float m_max = -INFINITY;
#pragma unroll
        for (int j = 0; j < NS; ++j) m_max = fmaxf(m_max, s_stat[((ts * NS + j) * 3 + 2) * AS_BM + row_in_tile]);
        sp[0] = m_max;
        sp[1] = l_all * ex2_f32(m_all - m_max);
      }
      // ---- O = sum_j w_j O_j / l : this stream normalises output columns [OC sub, OC sub + OC) ----
      constexpr int OC = AS_D / NS;   // 32
      float acc[OC];
#pragma unroll
      for (int j = 0; j < OC; ++j) acc[j] = 0.f;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        mbar_wait(&o_full[ts * NS + j], it & 1);
        tc_fence_after();
        uint32_t ov[OC];
        const uint32_t t_oj = tmem_base + (ts * NS + j) * 128 + lane_off + 64 + OC * sub;
        tmem_ld32(t_oj, ov);
        tmem_ld_wait();
        const float a = wgt[j] * inv;
#pragma unroll
        for (int i = 0; i < OC; ++i) acc[i] = fmaf(a, __uint_as_float(ov[i]), acc[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[ts]);   // the streams' accumulators may be overwritten
      // ---- staging tile = the tile's Q slot (every S product of the tile has retired: o_full is committed behind them) ----
      {
        const uint32_t srow = smem_u32(sQ + qslot * AS_QBYTES) + (uint32_t)row_in_tile * 128u;
#pragma unroll
        for (int j = 0; j < OC / 8; ++j) {
          const float* r = acc + 8 * j;
          st_shared_v4(srow + ((uint32_t)((OC / 8 * sub + j) ^ (row_in_tile & 7)) << 4), pack_bf16x2(r[0], r[1]),
                       pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, %1;" ::"r"(4 + ts), "n"(NS * 128) : "memory");
        if (storer) {
          tma_store_2d(&tmO, sQ + qslot * AS_QBYTES, head * AS_D, frame * p.Sq + tile * AS_BM);
          bulk_commit_group();
          pending_slot = qslot;
        }
      }
    }
    if (active && storer) bulk_wait_group<0>();   // the staging tiles must outlive the last stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int TQ>
static int launch_attn_streams(int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmQ, const CUtensorMap& tmK,
                               const CUtensorMap& tmV, const CUtensorMap& tmO, const AttnStreamParams& p) {
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(attn_d64_streams_kernel<TQ>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         smem_bytes));
    configured = true;
  }
  RALD_CHECK_CUDA(launch_pdl(attn_d64_streams_kernel<TQ>, dim3(grid), dim3(AS_THREADS), smem_bytes, stream, tmQ, tmK, tmV,
                             tmO, p));
  return 0;
}

bool attn_streams_enabled() {
  static const bool on = [] { const char* e = getenv("RALD_B200_ATTN_STREAMS"); return e == nullptr || e[0] != '0'; }();
  return on;
}

// Same contract as attn_d64_chunk (attn.cu) for Skv = 512.
int attn_d64_streams(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                     int64_t ldo, int frames, int heads, int Sq, int kv_frame_rows, float* stats, float scale,
                     cudaStream_t stream) {
  RALD_REQUIRE(frames > 0 && heads > 0 && Sq > 0 && Sq % AS_BM == 0, "attn streams: bad sizes");
  RALD_REQUIRE(kv_frame_rows >= AS_SKV, "attn streams: %d key rows per frame < 512", kv_frame_rows);
  RALD_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "attn streams: output not 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV, tmO;
  RALD_TRY(make_tmap_out(&tmO, O, (uint64_t)frames * Sq, (uint64_t)heads * AS_D, (uint64_t)ldo, false, AS_BM));
  RALD_TRY(make_tmap_2d_bf16(&tmQ, Q, (uint64_t)frames * Sq, (uint64_t)heads * AS_D, (uint64_t)ldq, AS_BM));
  const uint64_t kv_rows = (uint64_t)(frames - 1) * kv_frame_rows + AS_SKV;
  RALD_TRY(make_tmap_2d_bf16(&tmK, K, kv_rows, (uint64_t)heads * AS_D, (uint64_t)ldk, AS_CK));
  RALD_TRY(make_tmap_2d_bf16(&tmV, V, kv_rows, (uint64_t)heads * AS_D, (uint64_t)ldv, AS_CK));
  AttnStreamParams p;
  p.Sq = Sq;
  p.frames = frames;
  p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.kv_frame_rows = kv_frame_rows;
  p.stats = stats;
  p.dbg = g_attn_streams_dbg;
  static const int skew_env = [] { const char* e = getenv("RALD_B200_ATTN_SKEW"); return e ? atoi(e) : 1; }();
  p.skew = skew_env;
  const int q_tiles = Sq / AS_BM;
  const int sms = device_sm_count();
  // tiles per item: rounds x (four chunk rounds of an item + fill / merge); a round of two streams is shorter than a
  // round of four that share the exp2 unit
  int tq = 1;
  if (q_tiles % 2 == 0) {
    const long fh = (long)frames * heads;
    const double c1 = (double)((fh * q_tiles + sms - 1) / sms) * (4.0 * 0.7 + 0.7);
    const double c2 = (double)((fh * (q_tiles / 2) + sms - 1) / sms) * (4.0 * 1.0 + 0.7);
    if (c2 <= c1) tq = 2;
  }
  p.tq = tq;
  p.items_per_head = q_tiles / tq;
  p.num_items = frames * heads * p.items_per_head;
  const int smem_bytes = AS_QSLOTS * AS_QBYTES + 2 * AS_NCH * AS_CBYTES + AS_STREAMS * 3 * AS_BM * 4 + 512 + 1024;
  const int grid = p.num_items < sms ? p.num_items : sms;
  ProfScope prof(FAM_ATTN, stream, 4.0 * frames * heads * Sq * AS_SKV * AS_D);
  if (tq == 2) RALD_TRY(launch_attn_streams<2>(grid, smem_bytes, stream, tmQ, tmK, tmV, tmO, p));
  else RALD_TRY(launch_attn_streams<1>(grid, smem_bytes, stream, tmQ, tmK, tmV, tmO, p));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

extern "C" int rald_attn_streams_debug_buffer(unsigned long long* dev_buf) {
  rald::g_attn_streams_dbg = dev_buf;
  return 0;
}
