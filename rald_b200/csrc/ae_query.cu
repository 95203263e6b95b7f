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
//                   as the TMEM-resident A operand Qn (cols [0,256))
//   MMA (tcgen05) : S_i = Qn x K'_i^T for 4 chunks of 128 latents (A from TMEM, K' streamed by TMA through a
//                   smem ring), double-buffered in TMEM cols [256,384) / [384,512)
//   compute warps : online softmax over the chunks and the dot with v' in fp32 registers -> one logit per query.
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
constexpr int AQ_THREADS = 192;

struct AeQueryParams {
  const float* queries;   // [B, Q, 3]
  float* logits;          // [B, Q]
  const float* pe_bias;   // [512]
  const float* ln_g;      // [512] decoder_cross_attn.norm.weight
  const float* ln_b;      // [512]
  const float* vprime;    // [B, 512]
  const float* c0;        // [B]
  float freq[3][8];       // point_embed.basis block diagonal
  int B;
  int64_t Q;
  int tiles_per_frame;
  float scale_log2;       // log2(e) / sqrt(dim)
};

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
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_v + AQ_DIM);
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
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(feat_ready, 4);
    mbar_init(e_ready, 1);
    mbar_init(qn_ready, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&s_free[i], 4);
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
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: 4 W_pe tiles then 4 x 8 K' tiles per query tile =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int frame = tile / p.tiles_per_frame;
        for (int n = 0; n < 4; ++n) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], AQ_STAGE_BYTES);
          tma_load_2d(s_ring + s * AQ_STAGE_BYTES, &tmWpe, &full_bar[s], 0, n * 128);
          if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
        }
        for (int i = 0; i < 4; ++i) {
          for (int kb = 0; kb < 8; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1);
            mbar_arrive_expect_tx(&full_bar[s], AQ_STAGE_BYTES);
            tma_load_2d(s_ring + s * AQ_STAGE_BYTES, &tmKp, &full_bar[s], kb * 64, frame * AQ_DIM + i * 128);
            if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
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
          tc_commit(&empty_bar[s]);
          if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(e_ready);
        // ---- S_i = Qn x K'_i^T, Qn read from TMEM cols [0,256) ----
        mbar_wait(qn_ready, tile_ph);
        tc_fence_after();
        for (int i = 0; i < 4; ++i) {
          const int buf = i & 1;
          mbar_wait(&s_free[buf], free_ph[buf] ^ 1);
          free_ph[buf] ^= 1;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + 256 + buf * 128;
          for (int kb = 0; kb < 8; ++kb) {
            mbar_wait(&full_bar[s], ph);
            tc_fence_after();
            const uint64_t b_desc = make_sdesc_sw128(smem_u32(s_ring + s * AQ_STAGE_BYTES), 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_ts(d_tmem, tmem_base + kb * 32 + k * 8, b_desc + 2 * k, idesc, (kb | k) != 0);
            tc_commit(&empty_bar[s]);
            if (++s == AQ_STAGES) { s = 0; ph ^= 1; }
          }
          tc_commit(&s_ready[buf]);
        }
        tile_ph ^= 1;
      }
    }
  } else {
    // ===================== compute warps (128 threads, thread <-> query row <-> TMEM lane) =====================
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16);
    uint32_t tile_ph = 0;
    uint32_t ready_ph[2] = {0, 0};
    int cur_frame = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int frame = tile / p.tiles_per_frame;
      const int64_t q0 = static_cast<int64_t>(tile - frame * p.tiles_per_frame) * AQ_TILE;
      const int64_t qi = q0 + r;
      const bool q_ok = qi < p.Q;
      if (frame != cur_frame) {
        // all 128 compute threads finished the previous tile's reads of s_v before it is replaced
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = threadIdx.x - 64; i < AQ_DIM; i += 128) s_v[i] = p.vprime[(int64_t)frame * AQ_DIM + i];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        cur_frame = frame;
      }
      // ---- Fourier features (models_ae.py:128-137): [sin(p f) (24), cos(p f) (24), p (3)], zero padded to 64 ----
      float pt[3] = {0.f, 0.f, 0.f};
      if (q_ok) {
        const float* qp = p.queries + ((int64_t)frame * p.Q + qi) * 3;
        pt[0] = qp[0]; pt[1] = qp[1]; pt[2] = qp[2];
      }
      float f[AQ_FEAT];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float sn, cs;
          sincosf(pt[a] * p.freq[a][k], &sn, &cs);
          f[a * 8 + k] = sn;
          f[24 + a * 8 + k] = cs;
        }
      }
      f[48] = pt[0]; f[49] = pt[1]; f[50] = pt[2];
#pragma unroll
      for (int j = 51; j < AQ_FEAT; ++j) f[j] = 0.f;
      {
        uint8_t* row = s_feat + r * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = make_uint4(pack_bf16x2(f[8 * c + 0], f[8 * c + 1]), pack_bf16x2(f[8 * c + 2], f[8 * c + 3]),
                                     pack_bf16x2(f[8 * c + 4], f[8 * c + 5]), pack_bf16x2(f[8 * c + 6], f[8 * c + 7]));
          *reinterpret_cast<uint4*>(row + ((c ^ (r & 7)) << 4)) = v;  // 128-byte swizzle: chunk ^= row % 8
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(feat_ready);

      // ---- LayerNorm of e = E + bias (fp32) ; Qn (bf16 pairs) overwrites consumed E columns ----
      mbar_wait(e_ready, tile_ph);
      tc_fence_after();
      float sum = 0.f, sq = 0.f;
      for (int c = 0; c < AQ_DIM; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_lane + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float e = __uint_as_float(v[j]) + s_bias[c + j];
          sum += e;
          sq = fmaf(e, e, sq);
        }
      }
      const float mean = sum * (1.0f / AQ_DIM);
      const float var = fmaxf(sq * (1.0f / AQ_DIM) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      for (int c = 0; c < AQ_DIM; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_lane + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float e0 = (__uint_as_float(v[2 * j]) + s_bias[c + 2 * j] - mean) * rstd;
          const float e1 = (__uint_as_float(v[2 * j + 1]) + s_bias[c + 2 * j + 1] - mean) * rstd;
          pk[j] = pack_bf16x2(fmaf(e0, s_g[c + 2 * j], s_b[c + 2 * j]), fmaf(e1, s_g[c + 2 * j + 1], s_b[c + 2 * j + 1]));
        }
        tmem_st16(t_lane + (c >> 1), pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qn_ready);

      // ---- online softmax over 4 x 128 latents, fused with the dot against v' ----
      float m = -INFINITY, l = 0.f, acc = 0.f;
      for (int i = 0; i < 4; ++i) {
        const int buf = i & 1;
        mbar_wait(&s_ready[buf], ready_ph[buf]);
        ready_ph[buf] ^= 1;
        tc_fence_after();
        for (int c = 0; c < 128; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_lane + 256 + buf * 128 + c, v);
          tmem_ld_wait();
          float cm = __uint_as_float(v[0]);
#pragma unroll
          for (int j = 1; j < 32; ++j) cm = fmaxf(cm, __uint_as_float(v[j]));
          const float m_new = fmaxf(m, cm * p.scale_log2);
          const float corr = ex2f(m - m_new);  // first group: ex2(-inf) = 0
          l *= corr;
          acc *= corr;
          m = m_new;
          const float* vp = s_v + i * 128 + c;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float pj = ex2f(fmaf(__uint_as_float(v[j]), p.scale_log2, -m));
            l += pj;
            acc = fmaf(pj, vp[j], acc);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[buf]);
      }
      if (q_ok) p.logits[(int64_t)frame * p.Q + qi] = acc / l + p.c0[frame];
      tile_ph ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
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
  RALD_TRY(make_tmap_2d_bf16(&tmWpe, wpe_bf16, AQ_DIM, AQ_FEAT, AQ_FEAT, 128));
  RALD_TRY(make_tmap_2d_bf16(&tmKp, kprime_bf16, (uint64_t)B * AQ_DIM, AQ_DIM, AQ_DIM, 128));
  AeQueryParams p;
  p.queries = queries; p.logits = logits; p.pe_bias = pe_bias; p.ln_g = ln_g; p.ln_b = ln_b; p.vprime = vprime;
  p.c0 = c0;
  for (int a = 0; a < 3; ++a)
    for (int k = 0; k < 8; ++k) p.freq[a][k] = freq24[a * 8 + k];
  p.B = B; p.Q = Q;
  p.tiles_per_frame = (int)((Q + AQ_TILE - 1) / AQ_TILE);
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)dim);
  const int smem_bytes = AQ_TILE * AQ_FEAT * 2 + AQ_STAGES * AQ_STAGE_BYTES + 4 * AQ_DIM * 4 + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(ae_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = true;
  }
  const int64_t tiles = (int64_t)B * p.tiles_per_frame;
  const int sms = device_sm_count();
  const int grid = (int)(tiles < sms ? tiles : sms);
  ProfScope prof(FAM_AE_QUERY, stream, (double)B * Q * 577536.0);
  ae_query_kernel<<<grid, AQ_THREADS, smem_bytes, stream>>>(tmWpe, tmKp, p);
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald
