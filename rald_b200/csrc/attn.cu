// Fused softmax(Q K^T * scale) V for head_dim 64 and a key set that fits TMEM (Skv <= 512), tcgen05 + TMA.
//
// One CTA = 128 query rows of one (frame, head):
//   warp 0 lane 0 : TMA loads Q [128x64], K [Skv x 64], V [Skv x 64] (128B swizzle), then issues
//                   S = Q K^T   (tcgen05.mma SS, N = Skv in chunks of <=256, S fp32 in TMEM columns [0, Skv))
//                   O = P V     (tcgen05.mma TS: P read from TMEM, V as MN-major smem operand)
//   warps 1..4    : row r = TMEM lane r. sweep 1: row max over S; sweep 2: p = exp2((s - max) * scale*log2e),
//                   packed to bf16 pairs and written back INTO TMEM over the already-consumed S columns
//                   ([0, Skv/2)), row sum kept in fp32; after O is ready: O / sum -> bf16 -> global.
// The score matrix never leaves the SM (the reference materialises [B*h, Sq, Skv] fp32 in HBM:
// model/models_radar_generation.py:66-75, model/models_ae.py:91-104).
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int ATT_BM = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 160;

struct AttnParams {
  __nv_bfloat16* out;
  int64_t ldo;
  int Sq, Skv;
  float scale_log2;  // scale * log2(e)
  uint32_t tmem_cols;
  uint32_t o_col;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS)
attn_d64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int Skv = p.Skv;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_BM * ATT_D * 2;
  uint8_t* sV = sK + Skv * ATT_D * 2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Skv * ATT_D * 2);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_blk = blockIdx.x;
  const int head = blockIdx.y;
  const int frame = blockIdx.z;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      const int q_row0 = frame * p.Sq + q_blk * ATT_BM;
      const int kv_row0 = frame * Skv;
      mbar_arrive_expect_tx(bar_qk, (ATT_BM + Skv) * ATT_D * 2);
      tma_load_2d(sQ, &tmQ, bar_qk, head * ATT_D, q_row0);
      for (int r = 0; r < Skv; r += 256) tma_load_2d(sK + r * ATT_D * 2, &tmK, bar_qk, head * ATT_D, kv_row0 + r);
      mbar_arrive_expect_tx(bar_v, Skv * ATT_D * 2);
      for (int r = 0; r < Skv; r += 256) tma_load_2d(sV + r * ATT_D * 2, &tmV, bar_v, head * ATT_D, kv_row0 + r);

      // ---- S = Q K^T ----
      mbar_wait(bar_qk, 0);
      tc_fence_after();
      const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ), 16, 1024);
      for (int n0 = 0; n0 < Skv; n0 += 256) {
        const int n = (Skv - n0) < 256 ? (Skv - n0) : 256;
        const uint32_t idesc = make_idesc(FMT_BF16, ATT_BM, n, 0, 0);
        const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + n0 * ATT_D * 2), 16, 1024);
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) mma_f16_ss(tmem_base + n0, q_desc + 2 * k, k_desc + 2 * k, idesc, k != 0);
      }
      tc_commit(bar_s);

      // ---- O = P V (P in TMEM columns [0, Skv/2), 2 keys per 32-bit column) ----
      mbar_wait(bar_v, 0);
      mbar_wait(bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = make_idesc(FMT_BF16, ATT_BM, ATT_D, 0, 1);
      // V is [key][d] = MN-major B operand: 8-key groups are 1024 B apart (SBO); one 64-wide MN atom (LBO unused).
      const uint64_t v_desc = make_sdesc_sw128(smem_u32(sV), 1024, 1024);
      for (int k = 0; k < Skv / 16; ++k) {
        // 16 keys per MMA = 8 TMEM columns of P and 2048 B (=128 in >>4 units) of V
        mma_f16_ts(tmem_base + p.o_col, tmem_base + 8 * k, v_desc + 128 * k, idesc_o, k != 0);
      }
      tc_commit(bar_o);
    }
  } else {
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // sweep 1: row max
    float mx = -INFINITY;
    for (int c = 0; c < Skv; c += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float m_scaled = mx * p.scale_log2;
    // sweep 2: probabilities -> bf16 pairs written over consumed S columns
    float sum = 0.f;
    for (int c = 0; c < Skv; c += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + c, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * j]), p.scale_log2, -m_scaled));
        const float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * j + 1]), p.scale_log2, -m_scaled));
        const __nv_bfloat162 h = __floats2bfloat162_rn(e0, e1);
        sum += __low2float(h) + __high2float(h);
        pk[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      tmem_st16(t_lane + (c >> 1), pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p);

    // normalise and store O
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int64_t row = static_cast<int64_t>(frame) * p.Sq + q_blk * ATT_BM + row_in_tile;
    const bool row_ok = (q_blk * ATT_BM + row_in_tile) < p.Sq;
    uint4* dst = reinterpret_cast<uint4*>(p.out + row * p.ldo + head * ATT_D);
#pragma unroll
    for (int c = 0; c < ATT_D; c += 32) {
      uint32_t v[32];
      tmem_ld32(t_lane + p.o_col + c, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dst[(c >> 3) + j] = make_uint4(
              pack_bf16x2(__uint_as_float(v[8 * j + 0]) * inv, __uint_as_float(v[8 * j + 1]) * inv),
              pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv),
              pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv),
              pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int attn_d64(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
             int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, cudaStream_t stream) {
  RALD_REQUIRE(frames > 0 && heads > 0 && Sq > 0, "attn: bad sizes");
  RALD_REQUIRE(Skv >= 16 && Skv <= 512 && Skv % 32 == 0, "attn: Skv=%d must be a multiple of 32 in [32, 512]", Skv);
  RALD_REQUIRE(Sq % ATT_BM == 0, "attn: Sq=%d must be a multiple of 128", Sq);
  RALD_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(O) & 15) == 0, "attn: output not 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t kv_box = Skv < 256 ? Skv : 256;
  RALD_REQUIRE(Skv % kv_box == 0, "attn: Skv=%d must be <=256 or a multiple of 256", Skv);
  RALD_TRY(make_tmap_2d_bf16(&tmQ, Q, (uint64_t)frames * Sq, (uint64_t)heads * ATT_D, (uint64_t)ldq, ATT_BM));
  RALD_TRY(make_tmap_2d_bf16(&tmK, K, (uint64_t)frames * Skv, (uint64_t)heads * ATT_D, (uint64_t)ldk, kv_box));
  RALD_TRY(make_tmap_2d_bf16(&tmV, V, (uint64_t)frames * Skv, (uint64_t)heads * ATT_D, (uint64_t)ldv, kv_box));
  AttnParams p;
  p.out = reinterpret_cast<__nv_bfloat16*>(O);
  p.ldo = ldo;
  p.Sq = Sq;
  p.Skv = Skv;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o_col = (Skv + ATT_D <= 512) ? (uint32_t)Skv : (uint32_t)(Skv / 2);
  uint32_t need = p.o_col + ATT_D;
  if (need < (uint32_t)Skv) need = Skv;
  uint32_t cols = 32;
  while (cols < need) cols <<= 1;
  p.tmem_cols = cols;
  const int smem_bytes = (ATT_BM + 2 * Skv) * ATT_D * 2 + 1024 + 128;
  static int configured_bytes = 0;
  if (smem_bytes > configured_bytes) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(attn_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured_bytes = smem_bytes;
  }
  dim3 grid(Sq / ATT_BM, heads, frames);
  ProfScope prof(FAM_ATTN, stream, 4.0 * frames * heads * Sq * Skv * ATT_D);
  RALD_CHECK_CUDA(launch_pdl(attn_d64_kernel, grid, dim3(ATT_THREADS), smem_bytes, stream, tmQ, tmK, tmV, p));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald
