// the pieces of the reference's eval loop either side of the sampler + decoder
// that it runs in numpy / Python on the host.
//
//   refine_queries     aug_query_helper (datasets/utils/query_helper.py:3-43) + norm_points (utils/utils.py:78-104):
//                      the second-pass query set of `refine_query` (engine_generation.py:291-297) built on the device
//                      from the first pass's occupied points, so the points never visit the host in between.
//   chamfer            cal_metrics / chamfer_distance (utils/utils.py:116-142): per-point Python cKDTree.query loops
//                      -> brute-force nearest neighbour over shared-memory tiles, winner distance redone in fp64.
//   radar_cube_prep    Coloradar_dataset.process_radar_data (datasets/aligned_coloradar/Coloradar_dataset.py:432-475):
//                      clip / normalise intensity, mask / normalise doppler, bilinear (align_corners=True) upsample of
//                      the azimuth x elevation plane. The raw cube is 24 KB per frame, the upsampled one 2 MB: only
//                      the raw cube has to cross PCIe.
// All three are HBM / latency bound integer-and-float bookkeeping; no tensor-core work here.
#include <curand_kernel.h>

#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

// ------------------------------------------------------------------------------------------------------------
// refine_queries
// ------------------------------------------------------------------------------------------------------------
struct RefineParams {
  const float* pts;        // [cap, 3] occupied points of the first pass (inverse-normalised, polar)
  const int32_t* count;    // device scalar: number of valid rows of pts (may exceed cap: clamped)
  int64_t cap;
  int64_t aug_num;
  const int32_t* sel;      // [aug_num] injected draws (numpy parity mode) or null (device RNG)
  const int32_t* scl;      // [aug_num] injected integer scales in [1, aug_scale]
  const double* u;         // [aug_num, 3] injected uniforms in [0, 1)
  unsigned long long seed; // device RNG mode
  int aug_scale;
  double voxel[3];
  double lo[3], hi[3];     // pc_range
  float n_off[3], n_scale[3];  // norm_points constants (fp32, as numpy rounds the python scalars)
  float* out;              // [aug_num, 3] normalised refined queries
};

// Row i < N: the helper point itself. Row i >= N (generated row g = i - N):
//   p = float32( clip( double(pts[sel[g]]) + (u[g] * 2 - 1) * (voxel * scl[g]), lo, hi ) )      query_helper.py:33-41
// then norm_points in fp32: (p - offset) / scale                                                 utils/utils.py:95-98
// The injected arrays are indexed by g (exactly numpy's draw order: choice(N, G), choice(scales, G), rand(G, 3)).
__global__ void __launch_bounds__(256) refine_queries_kernel(const RefineParams p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.aug_num) return;
  int64_t N = *p.count;
  if (N > p.cap) N = p.cap;
  float x[3];
  if (i < N || N == 0) {
    // N == 0: the reference would raise inside np.random.choice(0, ...); rows are defined as the range centre here
    for (int a = 0; a < 3; ++a) x[a] = N == 0 ? (float)(0.5 * (p.lo[a] + p.hi[a])) : p.pts[i * 3 + a];
  } else {
    const int64_t g = i - N;
    int64_t s;
    double sc, uu[3];
    if (p.sel != nullptr) {
      s = p.sel[g];
      sc = (double)p.scl[g];
      for (int a = 0; a < 3; ++a) uu[a] = p.u[g * 3 + a];
    } else {
      curandStatePhilox4_32_10_t st;
      curand_init(p.seed, (unsigned long long)g, 0, &st);
      const uint4 r = curand4(&st);
      const uint4 r2 = curand4(&st);
      s = (int64_t)(((unsigned long long)r.x * (unsigned long long)N) >> 32);
      sc = (double)(1 + (int)(((unsigned long long)r.y * (unsigned long long)p.aug_scale) >> 32));
      uu[0] = ((double)r.z + 0.5) * (1.0 / 4294967296.0);
      uu[1] = ((double)r.w + 0.5) * (1.0 / 4294967296.0);
      uu[2] = ((double)r2.x + 0.5) * (1.0 / 4294967296.0);
    }
    if (s < 0) s = 0;
    if (s >= N) s = N - 1;
    for (int a = 0; a < 3; ++a) {
      const double bias = __dmul_rn(__dsub_rn(__dmul_rn(uu[a], 2.0), 1.0), __dmul_rn(p.voxel[a], sc));
      double v = __dadd_rn((double)p.pts[s * 3 + a], bias);
      v = fmin(fmax(v, p.lo[a]), p.hi[a]);
      x[a] = (float)v;
    }
  }
  for (int a = 0; a < 3; ++a) p.out[i * 3 + a] = __fdiv_rn(__fsub_rn(x[a], p.n_off[a]), p.n_scale[a]);
}

int refine_queries(const float* pts, const int32_t* count, int64_t cap, int64_t aug_num, const int32_t* sel,
                   const int32_t* scl, const double* u, unsigned long long seed, int aug_scale,
                   const double* voxel_host, const double* pc_range_host, const float* norm_host, float* out,
                   cudaStream_t stream) {
  RALD_REQUIRE(pts != nullptr && count != nullptr && out != nullptr, "refine_queries: null pointer");
  RALD_REQUIRE(cap > 0 && aug_num > 0 && aug_scale >= 1, "refine_queries: cap=%lld aug_num=%lld aug_scale=%d",
               (long long)cap, (long long)aug_num, aug_scale);
  RALD_REQUIRE((sel == nullptr) == (scl == nullptr) && (sel == nullptr) == (u == nullptr),
               "refine_queries: injected draws must be given all together or not at all");
  RefineParams p;
  p.pts = pts; p.count = count; p.cap = cap; p.aug_num = aug_num; p.sel = sel; p.scl = scl; p.u = u; p.seed = seed;
  p.aug_scale = aug_scale; p.out = out;
  for (int a = 0; a < 3; ++a) {
    p.voxel[a] = voxel_host[a];
    p.lo[a] = pc_range_host[a];
    p.hi[a] = pc_range_host[3 + a];
    p.n_scale[a] = norm_host[a];
    p.n_off[a] = norm_host[3 + a];
  }
  ProfScope prof(FAM_OTHER, stream, (double)aug_num * 24.0);
  refine_queries_kernel<<<(unsigned)((aug_num + 255) / 256), 256, 0, stream>>>(p);
  RALD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// chamfer
// ------------------------------------------------------------------------------------------------------------
constexpr int CH_THREADS = 256;
constexpr int CH_PER = 4;                       // source points per thread (each dst tile load is reused 4x)
constexpr int CH_SRC = CH_THREADS * CH_PER;     // source points per CTA
constexpr int CH_TILE = 2048;                   // destination points per shared-memory tile (24 KB)

// partial[dir][b][blk] = sum over the CTA's source points of the fp64 distance to their nearest destination point.
// dir 0: src = pred, dst = gt; dir 1: src = gt, dst = pred.
__global__ void __launch_bounds__(CH_THREADS)
chamfer_nn_kernel(const float* __restrict__ pred, const int32_t* __restrict__ pred_cnt, int64_t pred_cap,
                  const float* __restrict__ gt, const int32_t* __restrict__ gt_cnt, int64_t gt_cap, int gt_fixed,
                  int nblk, int B, double* __restrict__ partial) {
  const int b = blockIdx.y, dir = blockIdx.z;
  int64_t np_ = pred_cnt[b];
  if (np_ > pred_cap) np_ = pred_cap;
  int64_t ng = gt_cnt != nullptr ? (int64_t)gt_cnt[b] : (int64_t)gt_fixed;
  if (ng > gt_cap) ng = gt_cap;
  const float* src = dir == 0 ? pred + (int64_t)b * pred_cap * 3 : gt + (int64_t)b * gt_cap * 3;
  const float* dst = dir == 0 ? gt + (int64_t)b * gt_cap * 3 : pred + (int64_t)b * pred_cap * 3;
  const int64_t ns = dir == 0 ? np_ : ng;
  const int64_t nd = dir == 0 ? ng : np_;
  double* my_partial = partial + ((int64_t)dir * B + b) * nblk + blockIdx.x;
  const int64_t s0 = (int64_t)blockIdx.x * CH_SRC;
  if (s0 >= ns || nd == 0) {
    if (threadIdx.x == 0) *my_partial = 0.0;
    return;
  }
  __shared__ float sx[CH_TILE], sy[CH_TILE], sz[CH_TILE];
  __shared__ double s_red[CH_THREADS / 32];
  float px[CH_PER], py[CH_PER], pz[CH_PER], best[CH_PER];
  int besti[CH_PER];
#pragma unroll
  for (int k = 0; k < CH_PER; ++k) {
    const int64_t i = s0 + k * CH_THREADS + threadIdx.x;
    const int64_t ii = i < ns ? i : ns - 1;
    px[k] = src[ii * 3 + 0]; py[k] = src[ii * 3 + 1]; pz[k] = src[ii * 3 + 2];
    best[k] = __int_as_float(0x7f800000);
    besti[k] = 0;
  }
  for (int64_t t0 = 0; t0 < nd; t0 += CH_TILE) {
    const int n = (int)((nd - t0) < CH_TILE ? (nd - t0) : CH_TILE);
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += CH_THREADS) {
      sx[j] = dst[(t0 + j) * 3 + 0]; sy[j] = dst[(t0 + j) * 3 + 1]; sz[j] = dst[(t0 + j) * 3 + 2];
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const float qx = sx[j], qy = sy[j], qz = sz[j];
#pragma unroll
      for (int k = 0; k < CH_PER; ++k) {
        const float dx = px[k] - qx, dy = py[k] - qy, dz = pz[k] - qz;
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (d < best[k]) { best[k] = d; besti[k] = (int)(t0 + j); }
      }
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < CH_PER; ++k) {
    const int64_t i = s0 + k * CH_THREADS + threadIdx.x;
    if (i < ns) {
      const float* q = dst + (int64_t)besti[k] * 3;
      const double dx = (double)px[k] - (double)q[0], dy = (double)py[k] - (double)q[1],
                   dz = (double)pz[k] - (double)q[2];
      acc += sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < CH_THREADS / 32; ++w) t += s_red[w];
    *my_partial = t;
  }
}

// out[b] = {cd, mean NN distance pred -> gt, mean NN distance gt -> pred}; fixed summation order (deterministic).
// cd = +inf for an empty prediction (utils/utils.py:117-118) or an empty ground truth.
__global__ void __launch_bounds__(256)
chamfer_finish_kernel(const double* __restrict__ partial, const int32_t* __restrict__ pred_cnt, int64_t pred_cap,
                      const int32_t* __restrict__ gt_cnt, int64_t gt_cap, int gt_fixed, int nblk, int B,
                      double* __restrict__ out) {
  const int b = blockIdx.x;
  __shared__ double s_red[2][8];
  double t[2];
  for (int dir = 0; dir < 2; ++dir) {
    const double* pp = partial + ((int64_t)dir * B + b) * nblk;
    double a = 0.0;
    for (int i = threadIdx.x; i < nblk; i += 256) a += pp[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) s_red[dir][threadIdx.x >> 5] = a;
    t[dir] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int dir = 0; dir < 2; ++dir)
      for (int w = 0; w < 8; ++w) t[dir] += s_red[dir][w];
    int64_t np_ = pred_cnt[b];
    if (np_ > pred_cap) np_ = pred_cap;
    int64_t ng = gt_cnt != nullptr ? (int64_t)gt_cnt[b] : (int64_t)gt_fixed;
    if (ng > gt_cap) ng = gt_cap;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    if (np_ == 0 || ng == 0) {
      out[b * 3 + 0] = inf; out[b * 3 + 1] = inf; out[b * 3 + 2] = inf;
    } else {
      const double m0 = t[0] / (double)np_, m1 = t[1] / (double)ng;
      out[b * 3 + 0] = 0.5 * m1 + 0.5 * m0;
      out[b * 3 + 1] = m0;
      out[b * 3 + 2] = m1;
    }
  }
}

static int chamfer_nblk(int64_t pred_cap, int64_t gt_cap) {
  const int64_t m = pred_cap > gt_cap ? pred_cap : gt_cap;
  return (int)((m + CH_SRC - 1) / CH_SRC);
}

int chamfer(const float* pred, const int32_t* pred_cnt, int64_t pred_cap, const float* gt, const int32_t* gt_cnt,
            int64_t gt_cap, int gt_fixed, int B, double* out, double* ws, cudaStream_t stream) {
  RALD_REQUIRE(pred != nullptr && pred_cnt != nullptr && gt != nullptr && out != nullptr && ws != nullptr,
               "chamfer: null pointer");
  RALD_REQUIRE(B > 0 && B <= 65535 && pred_cap > 0 && gt_cap > 0 && pred_cap < (1ll << 31) && gt_cap < (1ll << 31),
               "chamfer: bad sizes B=%d pred_cap=%lld gt_cap=%lld", B, (long long)pred_cap, (long long)gt_cap);
  const int nblk = chamfer_nblk(pred_cap, gt_cap);
  ProfScope prof(FAM_OTHER, stream, (double)B * (double)pred_cap * (double)gt_cap * 2.0 * 8.0);
  chamfer_nn_kernel<<<dim3((unsigned)nblk, (unsigned)B, 2), CH_THREADS, 0, stream>>>(
      pred, pred_cnt, pred_cap, gt, gt_cnt, gt_cap, gt_fixed, nblk, B, ws);
  RALD_LAUNCHED();
  chamfer_finish_kernel<<<B, 256, 0, stream>>>(ws, pred_cnt, pred_cap, gt_cnt, gt_cap, gt_fixed, nblk, B, out);
  RALD_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// radar_cube_prep
// ------------------------------------------------------------------------------------------------------------
struct PrepParams {
  const float* raw;  // [B, R, A, E, C]
  float* out;        // [B, R, A_up, E_up, n_out]
  int B, R, A, E, C, A_up, E_up, n_out;
  int norm_intensity, norm_dopp;
  float max_intensity, max_dopp;
  float scale_a, scale_e;  // align_corners=True source-index scales, fp32 as ATen computes them
};

__device__ __forceinline__ float prep_channel(const PrepParams& p, const float* v, int ch) {
  // Coloradar_dataset.py:447-455: ch 0 = clip(I, 0, max_I) / max_I (zeros when norm_intensity is off);
  // ch 1 = doppler * valid-mask (last raw channel) (/ max_dopp when norm_dopp)
  if (ch == 0) {
    if (!p.norm_intensity) return 0.f;
    return __fdiv_rn(fminf(fmaxf(v[0], 0.f), p.max_intensity), p.max_intensity);
  }
  float d = __fmul_rn(v[1], v[p.C - 1]);
  if (p.norm_dopp) d = __fdiv_rn(d, p.max_dopp);
  return d;
}

// One thread per output voxel (b, r, a_up, e_up): both channels. Arithmetic of ATen's CPU upsample_bilinear2d as the
// reference's torch build evaluates it (F.interpolate(..., mode='bilinear', align_corners=True) on [1, R, A, E];
// pinned bit-for-bit by tests/golden/evalpost.npz): src = scale * dst in fp32, i0 = floor, lambda = src - i0,
// i1 = i0 + (i0 < in - 1), weights w_ij = w_a[i] * w_e[j] rounded to fp32, and
//   out = fma(x00, w00, fma(x01, w01, fma(x11, w11, x10 * w10))).
// An axis with in == out contributes (i0 = i1 = dst, weights 1 and 0).
__global__ void __launch_bounds__(256) radar_cube_prep_kernel(const PrepParams p) {
  const int64_t total = (int64_t)p.B * p.R * p.A_up * p.E_up;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = (int)(i % p.E_up);
  const int a = (int)((i / p.E_up) % p.A_up);
  const int64_t br = i / ((int64_t)p.E_up * p.A_up);
  const float* cube = p.raw + br * p.A * p.E * p.C;
  int a0, a1, e0, e1;
  float la, le;
  if (p.A_up == p.A) { a0 = a1 = a; la = 0.f; }
  else {
    const float s = __fmul_rn(p.scale_a, (float)a);
    a0 = min((int)floorf(s), p.A - 1);
    la = fminf(fmaxf(__fsub_rn(s, (float)a0), 0.f), 1.f);
    a1 = a0 + (a0 < p.A - 1 ? 1 : 0);
  }
  if (p.E_up == p.E) { e0 = e1 = e; le = 0.f; }
  else {
    const float s = __fmul_rn(p.scale_e, (float)e);
    e0 = min((int)floorf(s), p.E - 1);
    le = fminf(fmaxf(__fsub_rn(s, (float)e0), 0.f), 1.f);
    e1 = e0 + (e0 < p.E - 1 ? 1 : 0);
  }
  const float wa0 = __fsub_rn(1.f, la), we0 = __fsub_rn(1.f, le);
  const float* v00 = cube + ((int64_t)a0 * p.E + e0) * p.C;
  const float* v01 = cube + ((int64_t)a0 * p.E + e1) * p.C;
  const float* v10 = cube + ((int64_t)a1 * p.E + e0) * p.C;
  const float* v11 = cube + ((int64_t)a1 * p.E + e1) * p.C;
  float* o = p.out + i * p.n_out;
  for (int ch = 0; ch < p.n_out; ++ch) {
    const float x00 = prep_channel(p, v00, ch), x01 = prep_channel(p, v01, ch);
    const float x10 = prep_channel(p, v10, ch), x11 = prep_channel(p, v11, ch);
    const float w00 = __fmul_rn(wa0, we0), w01 = __fmul_rn(wa0, le), w10 = __fmul_rn(la, we0), w11 = __fmul_rn(la, le);
    o[ch] = __fmaf_rn(x00, w00, __fmaf_rn(x01, w01, __fmaf_rn(x11, w11, __fmul_rn(x10, w10))));
  }
}

int radar_cube_prep(const float* raw, int B, int R, int A, int E, int C, int A_up, int E_up, int n_out,
                    int norm_intensity, float max_intensity, int norm_dopp, float max_dopp, float* out,
                    cudaStream_t stream) {
  RALD_REQUIRE(raw != nullptr && out != nullptr, "radar_cube_prep: null pointer");
  RALD_REQUIRE(B > 0 && R > 0 && A > 0 && E > 0 && C >= 2 && A_up >= 1 && E_up >= 1 && (n_out == 1 || n_out == 2),
               "radar_cube_prep: bad shape B=%d R=%d A=%d E=%d C=%d -> A_up=%d E_up=%d channels=%d", B, R, A, E, C, A_up,
               E_up, n_out);
  RALD_REQUIRE(!norm_intensity || max_intensity > 0.f, "radar_cube_prep: max_intensity=%f", (double)max_intensity);
  PrepParams p;
  p.raw = raw; p.out = out; p.B = B; p.R = R; p.A = A; p.E = E; p.C = C; p.A_up = A_up; p.E_up = E_up; p.n_out = n_out;
  p.norm_intensity = norm_intensity; p.norm_dopp = norm_dopp; p.max_intensity = max_intensity; p.max_dopp = max_dopp;
  // area_pixel_compute_scale<float>(in, out, align_corners=true): out > 1 ? (float)(in - 1) / (out - 1) : 0
  p.scale_a = A_up > 1 ? (float)(A - 1) / (float)(A_up - 1) : 0.f;
  p.scale_e = E_up > 1 ? (float)(E - 1) / (float)(E_up - 1) : 0.f;
  const int64_t total = (int64_t)B * R * A_up * E_up;
  ProfScope prof(FAM_OTHER, stream, (double)total * 4.0 * n_out);
  radar_cube_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p);
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

extern "C" {

int rald_refine_queries(const float* points, const int32_t* count, int64_t cap, int64_t aug_num, const int32_t* sel,
                        const int32_t* scales, const double* uniforms, uint64_t seed, int aug_scale,
                        const double* voxel_size_host, const double* pc_range_host, const float* norm_scale_offset_host,
                        float* out, void* stream) {
  return rald::refine_queries(points, count, cap, aug_num, sel, scales, uniforms, seed, aug_scale, voxel_size_host,
                              pc_range_host, norm_scale_offset_host, out, static_cast<cudaStream_t>(stream));
}

int rald_chamfer(const float* pred, const int32_t* pred_counts, int64_t pred_cap, const float* gt,
                 const int32_t* gt_counts, int64_t gt_cap, int gt_fixed, int B, double* out, double* ws, void* stream) {
  return rald::chamfer(pred, pred_counts, pred_cap, gt, gt_counts, gt_cap, gt_fixed, B, out, ws,
                       static_cast<cudaStream_t>(stream));
}

int64_t rald_chamfer_ws_elems(int B, int64_t pred_cap, int64_t gt_cap) {
  return 2ll * B * rald::chamfer_nblk(pred_cap, gt_cap);
}

int rald_radar_cube_prep(const float* raw, int B, int R, int A, int E, int C, int A_up, int E_up, int channels_out,
                         int norm_intensity, float max_intensity, int norm_dopp, float max_dopp, float* out,
                         void* stream) {
  return rald::radar_cube_prep(raw, B, R, A, E, C, A_up, E_up, channels_out, norm_intensity, max_intensity, norm_dopp,
                               max_dopp, out, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
