// Training- and AE-evaluation-loop helpers of SURVEY.md §8(f) rows 3 and 4 that need no autograd.
//
//   update_ema (engine_generation.py:29-39): for every (target, source) parameter pair
//       targ.detach().mul_(rate).add_(src, alpha=1 - rate)
//   The reference issues two elementwise launches per parameter (2 x 637 for the default denoiser, each far too small
//   to fill the machine and reading `targ` twice). Here the whole parameter list is ONE launch: a device table of
//   (target pointer, source pointer, element count, first chunk) per tensor, persistent CTAs walking 4096-element
//   chunks, 128-bit accesses. HBM-bound: 12 B per parameter element (target read + written, source read).
//   Arithmetic = ATen's, bit for bit: t1 = rn(targ * rate_f32); out = fma(src, alpha_f32, t1) (both the vectorised
//   CPU kernel and the CUDA kernel of torch contract `a + alpha * b` into one fma; checked against torch CPU in
//   tests/test_cpu_oracle_golden.py and against the unmodified reference function in tests/golden/ema.npz).
#include "../../include/rald_b200.h"

#include "host.cuh"
#include "kernels.h"

namespace rald {

constexpr int EMA_THREADS = 256;
constexpr int EMA_CHUNK = 4096;  // elements per chunk = 256 threads x 4 float4

// table (int64, device): [0,n) target pointers, [n,2n) source pointers, [2n,3n) element counts,
// [3n, 4n+1) first chunk of each tensor (exclusive prefix of ceil(count / EMA_CHUNK); last entry = total chunks)
__global__ void __launch_bounds__(EMA_THREADS)
ema_update_kernel(const int64_t* __restrict__ table, int n, float rate, float alpha) {
  const int64_t* tgt_p = table;
  const int64_t* src_p = table + n;
  const int64_t* cnt_p = table + 2 * (int64_t)n;
  const int64_t* first = table + 3 * (int64_t)n;
  const int64_t total = first[n];
  for (int64_t c = blockIdx.x; c < total; c += gridDim.x) {
    // tensor that owns chunk c: last i with first[i] <= c (empty tensors own no chunk and are skipped by the search)
    int lo = 0, hi = n;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (first[mid] <= c) lo = mid; else hi = mid;
    }
    float* __restrict__ t = reinterpret_cast<float*>(tgt_p[lo]);
    const float* __restrict__ s = reinterpret_cast<const float*>(src_p[lo]);
    const int64_t cnt = cnt_p[lo];
    const int64_t e0 = (c - first[lo]) * EMA_CHUNK;
    const int64_t e1 = min(e0 + (int64_t)EMA_CHUNK, cnt);
    const bool vec = (((uintptr_t)t | (uintptr_t)s) & 15) == 0 && e1 - e0 == EMA_CHUNK;
    if (vec) {
      float4* t4 = reinterpret_cast<float4*>(t + e0);
      const float4* s4 = reinterpret_cast<const float4*>(s + e0);
      float4 a[4], b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j] = t4[threadIdx.x + j * EMA_THREADS];
        b[j] = __ldg(s4 + threadIdx.x + j * EMA_THREADS);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[j].x = __fmaf_rn(b[j].x, alpha, __fmul_rn(a[j].x, rate));
        a[j].y = __fmaf_rn(b[j].y, alpha, __fmul_rn(a[j].y, rate));
        a[j].z = __fmaf_rn(b[j].z, alpha, __fmul_rn(a[j].z, rate));
        a[j].w = __fmaf_rn(b[j].w, alpha, __fmul_rn(a[j].w, rate));
        t4[threadIdx.x + j * EMA_THREADS] = a[j];
      }
    } else {
      for (int64_t e = e0 + threadIdx.x; e < e1; e += EMA_THREADS)
        t[e] = __fmaf_rn(s[e], alpha, __fmul_rn(t[e], rate));
    }
  }
}

int ema_update(const int64_t* table_dev, int n, int64_t total_chunks, int64_t total_elems, float rate, float alpha,
               cudaStream_t stream) {
  RALD_REQUIRE(table_dev != nullptr, "ema_update: null table");
  RALD_REQUIRE(n > 0 && total_chunks >= 0 && total_elems >= 0, "ema_update: bad sizes n=%d chunks=%lld elems=%lld", n,
               (long long)total_chunks, (long long)total_elems);
  RALD_REQUIRE(rate == rate && alpha == alpha, "ema_update: NaN rate");
  if (total_chunks == 0) return 0;
  const int64_t max_grid = (int64_t)device_sm_count() * 8;  // 8 CTAs of 256 threads per SM, persistent
  const unsigned grid = (unsigned)(total_chunks < max_grid ? total_chunks : max_grid);
  ProfScope prof(FAM_OTHER, stream, (double)total_elems * 12.0);
  ema_update_kernel<<<grid, EMA_THREADS, 0, stream>>>(table_dev, n, rate, alpha);
  RALD_LAUNCHED();
  return 0;
}

// ---- occupancy accuracy / IoU of the AE evaluation loops (engine_generation.py:376-385 in cache_latents, the same
// lines in engine_ae.py's evaluate): pred = logits >= threshold; accuracy = mean(pred == labels);
// iou = sum(pred * labels) / count(pred + labels > 0) + 1e-5. Labels are the dataset's 0 / 1 occupancy labels, so all
// three sums are integer counts (exact; the reference's fp32 sums of ones are exact up to 2^24 queries as well).
constexpr int IOU_THREADS = 256;
constexpr int IOU_PER_THREAD = 16;

__global__ void __launch_bounds__(IOU_THREADS)
occ_iou_count_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int64_t Q, float thr,
                     int32_t* __restrict__ counts) {
  const int b = blockIdx.y;
  const float* lg = logits + (int64_t)b * Q;
  const float* lb = labels + (int64_t)b * Q;
  const int64_t base = (int64_t)blockIdx.x * (IOU_THREADS * IOU_PER_THREAD);
  int eq = 0, inter = 0, uni = 0;
#pragma unroll 4
  for (int j = 0; j < IOU_PER_THREAD; ++j) {
    const int64_t q = base + (int64_t)j * IOU_THREADS + threadIdx.x;
    if (q < Q) {
      const float pred = lg[q] >= thr ? 1.f : 0.f;
      const float l = lb[q];
      eq += pred == l ? 1 : 0;
      inter += (pred * l) != 0.f ? 1 : 0;
      uni += (pred + l) > 0.f ? 1 : 0;
    }
  }
  eq = __reduce_add_sync(0xffffffffu, eq);
  inter = __reduce_add_sync(0xffffffffu, inter);
  uni = __reduce_add_sync(0xffffffffu, uni);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(counts + b * 3 + 0, eq);
    atomicAdd(counts + b * 3 + 1, inter);
    atomicAdd(counts + b * 3 + 2, uni);
  }
}

__global__ void occ_iou_finish_kernel(const int32_t* __restrict__ counts, int B, int64_t Q, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  out[b * 2 + 0] = __fdiv_rn((float)counts[b * 3 + 0], (float)Q);
  // fp32: intersection * 1.0 / union + 1e-5 (0 / 0 -> NaN as in the reference)
  out[b * 2 + 1] = __fadd_rn(__fdiv_rn((float)counts[b * 3 + 1], (float)counts[b * 3 + 2]), 1e-5f);
}

int occupancy_iou(const float* logits, const float* labels, int B, int64_t Q, float thr, float* out, int32_t* ws,
                  cudaStream_t stream) {
  RALD_REQUIRE(logits != nullptr && labels != nullptr && out != nullptr && ws != nullptr, "occupancy_iou: null pointer");
  RALD_REQUIRE(B > 0 && B <= 65535 && Q > 0 && Q < (1ll << 31), "occupancy_iou: bad sizes B=%d Q=%lld", B, (long long)Q);
  RALD_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(int32_t) * 3 * B, stream));
  const int64_t per = IOU_THREADS * IOU_PER_THREAD;
  ProfScope prof(FAM_OTHER, stream, (double)B * Q * 8.0);
  occ_iou_count_kernel<<<dim3((unsigned)((Q + per - 1) / per), (unsigned)B), IOU_THREADS, 0, stream>>>(logits, labels, Q,
                                                                                                      thr, ws);
  RALD_LAUNCHED();
  occ_iou_finish_kernel<<<(B + 127) / 128, 128, 0, stream>>>(ws, B, Q, out);
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

extern "C" {

int rald_occupancy_iou(const float* logits, const float* labels, int B, int64_t Q, float threshold, float* out,
                       int32_t* ws, void* stream) {
  return rald::occupancy_iou(logits, labels, B, Q, threshold, out, ws, static_cast<cudaStream_t>(stream));
}

int rald_ema_update(const int64_t* table_dev, int n_tensors, int64_t total_chunks, int64_t total_elems, float rate,
                    float one_minus_rate, void* stream) {
  return rald::ema_update(table_dev, n_tensors, total_chunks, total_elems, rate, one_minus_rate,
                          static_cast<cudaStream_t>(stream));
}

int rald_ema_chunk_elems(void) { return rald::EMA_CHUNK; }

}  // extern "C"
