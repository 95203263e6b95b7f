// 3x3x3 Conv3d as an implicit GEMM on tcgen05 for channels-last activations (radar encoder,
// model/models_radar_encoder.py:63-72 ResnetBlock convs, :34-41 stride-2 Downsample, :208-214 conv_out).
//
//   out[n, d, h, w, :] = bias + sum_{tap, ci} W[:, tap, ci] * x[n, s*d + kd - p, s*h + kh - p, s*w + kw - p, ci]
//
// GEMM view: M = output voxels (tile = 128 voxels = a (bn, bd, bh, bw) box of the output grid), N = Cout,
// K = 27 taps x Cin. The A tile of every (tap, 64-channel chunk) is ONE 5-D TMA box load of the bf16 activation
// tensor [N, D, H, W, C] at shifted coordinates; out-of-bounds voxels are zero-filled by TMA, which implements
// both the symmetric padding of the stride-1 convs and the high-side (0,1) padding of Downsample. No im2col
// buffer ever exists. Stride 2 uses eight "parity" tensor maps (base offset by (pd, ph, pw), doubled strides),
// so each tap is still a dense box load.
//
// Pipeline / roles are those of gemm.cu: warp 0 TMA producer, warp 1 tcgen05.mma issuer, warps 2..5 epilogue
// (bias + optional fp32 residual, fp32 channels-last output).
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int CV_BM = 128;
constexpr int CV_THREADS = 192;

struct ConvMaps {
  CUtensorMap a[8];  // stride 1: a[0]; stride 2: a[(kd&1)*4 + (kh&1)*2 + (kw&1)]
};

struct ConvParams {
  float* out;          // [B, Do, Ho, Wo, Cout]
  const float* bias;   // [Cout] (padded to the N tile)
  const float* resid;  // same shape as out, or null
  int B, Do, Ho, Wo, Cin, Cout;
  int bn, bd, bh, bw;  // output box of one M tile (product 128)
  int tiles_n, tiles_d, tiles_h, tiles_w;
  int stride;
  int num_n_blks;
};

template <int BN>
struct ConvCfg {
  static constexpr int A_BYTES = CV_BM * 64 * 2;
  static constexpr int B_BYTES = BN * 64 * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // CV_G k-blocks (tap x 64 input channels) share one ring slot = one full / empty barrier pair: an mbarrier probe
  // between groups of 4 MMAs costs the issuing thread ~260 cycles (tools/micro/mma_rate.cu), more than the 192 cycles
  // of tensor work of a k-block at N = 64 — with one probe per k-block the level-0 convolutions ran at ~500 cycles
  // per k-block. 27 taps x Cin / 64 chunks is always a multiple of 3.
  static constexpr int G = 3;
  static constexpr int STAGES_RAW = (216 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = (STAGES_RAW > 9 ? 9 : STAGES_RAW) / G * G;
  static constexpr int SLOTS = STAGES / G;
  static constexpr int SLOT_BYTES = G * STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(SLOTS >= 2, "ring too shallow");
};

template <int BN>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3d_kernel(const __grid_constant__ ConvMaps maps, const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
  using Cfg = ConvCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = p.tiles_n * p.tiles_d * p.tiles_h * p.tiles_w;
  const int num_tiles = m_tiles * p.num_n_blks;
  const int c_chunks = p.Cin / 64;
  const int num_kb = 27 * c_chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&maps.a[0]);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % p.num_n_blks;
        int mt = tile / p.num_n_blks;
        const int tw = mt % p.tiles_w; mt /= p.tiles_w;
        const int th = mt % p.tiles_h; mt /= p.tiles_h;
        const int td = mt % p.tiles_d; mt /= p.tiles_d;
        const int tn = mt;
        const int w0 = tw * p.bw, h0 = th * p.bh, d0 = td * p.bd, n0 = tn * p.bn;
        for (int kb0 = 0; kb0 < num_kb; kb0 += Cfg::G) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], Cfg::SLOT_BYTES);
#pragma unroll
          for (int g = 0; g < Cfg::G; ++g) {
            const int kb = kb0 + g;
            const int tap = kb / c_chunks, cc = kb - tap * c_chunks;
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            int mi, od, oh, ow;
            if (p.stride == 1) {
              mi = 0; od = kd - 1; oh = kh - 1; ow = kw - 1;
            } else {
              mi = (kd & 1) * 4 + (kh & 1) * 2 + (kw & 1);
              od = kd >> 1; oh = kh >> 1; ow = kw >> 1;
            }
            uint8_t* sa = smem + s * Cfg::SLOT_BYTES + g * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            tma_load_5d(sa, &maps.a[mi], &full_bar[s], cc * 64, w0 + ow, h0 + oh, d0 + od, n0);
            tma_load_2d(sb, &tmW, &full_bar[s], tap * p.Cin + cc * 64, n_blk * BN);
          }
          if (++s == Cfg::SLOTS) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {   // one elected lane: the compiler keeps descriptors / coordinates in uniform registers (no ELECT / R2UR.BROADCAST loop per tcgen05 / TMA instruction)
      constexpr uint32_t idesc = make_idesc(FMT_BF16, CV_BM, BN, 0, 0);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb0 = 0; kb0 < num_kb; kb0 += Cfg::G) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < Cfg::G; ++g) {
            const uint32_t sa = smem_u32(smem + s * Cfg::SLOT_BYTES + g * Cfg::STAGE_BYTES);
            const uint64_t a_desc = make_sdesc_sw128(sa, 16, 1024);
            const uint64_t b_desc = make_sdesc_sw128(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              mma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, ((kb0 + g) | k) != 0 ? 1u : 0u);
          }
          tc_commit(&empty_bar[s]);
          if (++s == Cfg::SLOTS) { s = 0; ph ^= 1; }
        }
        tc_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;  // row inside the tile
    // tile-local voxel coordinates of this row (w fastest)
    const int lw = m % p.bw;
    const int lh = (m / p.bw) % p.bh;
    const int ld = (m / (p.bw * p.bh)) % p.bd;
    const int ln = m / (p.bw * p.bh * p.bd);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile % p.num_n_blks;
      int mt = tile / p.num_n_blks;
      const int tw = mt % p.tiles_w; mt /= p.tiles_w;
      const int th = mt % p.tiles_h; mt /= p.tiles_h;
      const int td = mt % p.tiles_d; mt /= p.tiles_d;
      const int tn = mt;
      const int w = tw * p.bw + lw, h = th * p.bh + lh, d = td * p.bd + ld, n = tn * p.bn + ln;
      const bool row_ok = (n < p.B) && (d < p.Do) && (h < p.Ho) && (w < p.Wo);
      const int64_t vox = (((int64_t)n * p.Do + d) * p.Ho + h) * p.Wo + w;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= p.Cout) break;
        uint32_t v[32];
        tmem_ld32(t_row + c * 32, v);
        tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + b.x);
          v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + b.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + b.z);
          v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + b.w);
        }
        if (row_ok) {
          // Cout may be smaller than the 32-column group (conv_out: 16 channels): store only valid columns
          const int ncol = (p.Cout - col0) < 32 ? (p.Cout - col0) : 32;
          float* dst = p.out + vox * p.Cout + col0;
          if (p.resid != nullptr) {
            const float* rs = p.resid + vox * p.Cout + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (4 * j < ncol) {
                const float4 r = *reinterpret_cast<const float4*>(rs + 4 * j);
                v[4 * j + 0] = __float_as_uint(__uint_as_float(v[4 * j + 0]) + r.x);
                v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + r.y);
                v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + r.z);
                v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + r.w);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (4 * j < ncol)
              *reinterpret_cast<uint4*>(dst + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN>
static int launch_conv(const ConvMaps& maps, const CUtensorMap& tmW, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  auto kern = conv3d_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int num_tiles = p.tiles_n * p.tiles_d * p.tiles_h * p.tiles_w * p.num_n_blks;
  const int sms = device_sm_count();
  const int grid = num_tiles < sms ? num_tiles : sms;
  ProfScope prof(FAM_CONV3D, stream, 2.0 * p.B * p.Do * p.Ho * p.Wo * (double)p.Cout * 27.0 * p.Cin);
  kern<<<grid, CV_THREADS, Cfg::SMEM_BYTES, stream>>>(maps, tmW, p);
  RALD_LAUNCHED();
  return 0;
}

int conv3d_cl(const void* x_bf16, const void* w_packed, int w_rows, const float* bias, const float* resid, float* out,
              int B, int D, int H, int W, int Cin, int Cout, int stride, cudaStream_t stream) {
  RALD_REQUIRE(stride == 1 || stride == 2, "conv3d: stride %d", stride);
  RALD_REQUIRE(Cin % 64 == 0, "conv3d: Cin=%d must be a multiple of 64", Cin);
  RALD_REQUIRE(Cout % 4 == 0, "conv3d: Cout=%d must be a multiple of 4", Cout);
  RALD_REQUIRE(stride == 1 || (D % 2 == 0 && H % 2 == 0 && W % 2 == 0), "conv3d: stride 2 needs even D/H/W");
  const int Do = D / stride, Ho = H / stride, Wo = W / stride;
  auto pow2_le = [](int v, int cap) { int r = 1; while (r * 2 <= v && r * 2 <= cap) r *= 2; return r; };
  ConvParams p;
  p.bw = pow2_le(Wo, 128);
  p.bh = pow2_le(Ho, 128 / p.bw);
  p.bd = pow2_le(Do, 128 / (p.bw * p.bh));
  p.bn = 128 / (p.bw * p.bh * p.bd);
  RALD_REQUIRE(Wo % p.bw == 0 && Ho % p.bh == 0 && Do % p.bd == 0,
               "conv3d: output grid %dx%dx%d is not tileable by %dx%dx%d boxes", Do, Ho, Wo, p.bd, p.bh, p.bw);
  p.tiles_w = Wo / p.bw; p.tiles_h = Ho / p.bh; p.tiles_d = Do / p.bd; p.tiles_n = (B + p.bn - 1) / p.bn;
  p.out = out; p.bias = bias; p.resid = resid;
  p.B = B; p.Do = Do; p.Ho = Ho; p.Wo = Wo; p.Cin = Cin; p.Cout = Cout; p.stride = stride;
  int bn_tile = Cout >= 128 ? 128 : (Cout >= 64 ? 64 : 32);
  RALD_REQUIRE(w_rows % bn_tile == 0 && w_rows >= Cout, "conv3d: packed weight rows %d must cover Cout padded to %d",
               w_rows, bn_tile);
  p.num_n_blks = (Cout + bn_tile - 1) / bn_tile;

  ConvMaps maps;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x_bf16);
  if (stride == 1) {
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)B};
    uint64_t str[4] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2, (uint64_t)D * H * W * Cin * 2};
    uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
    RALD_TRY(make_tmap_nd_bf16(&maps.a[0], xb, 5, dims, str, box, nullptr));
    for (int i = 1; i < 8; ++i) maps.a[i] = maps.a[0];
  } else {
    for (int i = 0; i < 8; ++i) {
      const int pd = (i >> 2) & 1, ph = (i >> 1) & 1, pw = i & 1;
      const __nv_bfloat16* base = xb + (((int64_t)pd * H + ph) * W + pw) * Cin;
      // sub-lattice of voxels with parity (pd, ph, pw): extents (D-pd+1)/2 etc., doubled strides
      uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)((W - pw + 1) / 2), (uint64_t)((H - ph + 1) / 2),
                          (uint64_t)((D - pd + 1) / 2), (uint64_t)B};
      uint64_t str[4] = {(uint64_t)2 * Cin * 2, (uint64_t)2 * W * Cin * 2, (uint64_t)2 * H * W * Cin * 2,
                         (uint64_t)D * H * W * Cin * 2};
      uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
      RALD_TRY(make_tmap_nd_bf16(&maps.a[i], base, 5, dims, str, box, nullptr));
    }
  }
  CUtensorMap tmW;
  RALD_TRY(make_tmap_2d_bf16(&tmW, w_packed, (uint64_t)w_rows, (uint64_t)27 * Cin, (uint64_t)27 * Cin, (uint32_t)bn_tile));
  switch (bn_tile) {
    case 32: return launch_conv<32>(maps, tmW, p, stream);
    case 64: return launch_conv<64>(maps, tmW, p, stream);
    default: return launch_conv<128>(maps, tmW, p, stream);
  }
}

}  // namespace rald
