// Device-side post-processing of decoded occupancy logits (engine_generation.py:283-289, 313-315 in the reference,
// where it is numpy on the host after a synchronous .cpu() of all logits):
//   occupied = logits > threshold                               np.where(output_np > 0)            :285
//   points   = queries[occupied] * scale + offset               inverse_norm_points                utils/utils.py:50-76
//   points   = polar (r, az deg, el deg) -> cartesian           polar2cartesian   dataset_preprocessor/lidar.py:57-63
// Stable stream compaction (query order is preserved, so the result is identical to np.where + gather):
//   pass 1  per-1024-query block counts; pass 2 exclusive scan per frame; pass 3 ordered scatter.
// Only the occupied points (a few % of the queries) then have to leave the GPU. HBM-bound: 4 B/query (pass 1)
// + 4 B/query + 12 B/occupied query read, 12 (+4) B/occupied query written (pass 3).
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

constexpr int PP_THREADS = 256;
constexpr int PP_PER_THREAD = 4;
constexpr int PP_BLOCK = PP_THREADS * PP_PER_THREAD;  // queries per block

__global__ void __launch_bounds__(PP_THREADS)
occ_count_kernel(const float* __restrict__ logits, int64_t Q, float thr, int nblk, int32_t* __restrict__ block_counts) {
  const int b = blockIdx.y;
  const int64_t q0 = (int64_t)blockIdx.x * PP_BLOCK + threadIdx.x * PP_PER_THREAD;
  const float* lg = logits + (int64_t)b * Q;
  int c = 0;
#pragma unroll
  for (int j = 0; j < PP_PER_THREAD; ++j)
    if (q0 + j < Q) c += lg[q0 + j] > thr ? 1 : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int s_c[PP_THREADS / 32];
  if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < PP_THREADS / 32; ++i) t += s_c[i];
    block_counts[(int64_t)b * nblk + blockIdx.x] = t;
  }
}

// In-place exclusive scan of block_counts[b][0..nblk); counts[b] = total. One CTA per frame.
__global__ void __launch_bounds__(1024)
occ_scan_kernel(int32_t* __restrict__ block_counts, int nblk, int32_t* __restrict__ counts) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int b = blockIdx.x;
  int32_t* bc = block_counts + (int64_t)b * nblk;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? bc[i] : 0;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if ((threadIdx.x & 31) >= o) incl += n;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = s_warp[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += n;
      }
      s_warp[threadIdx.x] = w;  // inclusive over warps
    }
    __syncthreads();
    const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
    const int carry = s_carry;
    if (i < nblk) bc[i] = carry + warp_off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[b] = s_carry;
}

struct OccParams {
  const float* logits;
  const float* queries;
  const int32_t* block_offsets;
  float* points;
  int32_t* index;
  int64_t Q, cap;
  int nblk;
  float thr;
  float scale[3], offset[3];
  int polar2cart;
};

__global__ void __launch_bounds__(PP_THREADS)
occ_scatter_kernel(const OccParams p) {
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * PP_BLOCK + threadIdx.x * PP_PER_THREAD;
  const float* lg = p.logits + (int64_t)b * p.Q;
  bool occ[PP_PER_THREAD];
  int c = 0;
#pragma unroll
  for (int j = 0; j < PP_PER_THREAD; ++j) {
    occ[j] = (q0 + j < p.Q) && (lg[q0 + j] > p.thr);
    c += occ[j] ? 1 : 0;
  }
  // exclusive prefix of c inside the block (thread order = query order)
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  __shared__ int s_w[PP_THREADS / 32];
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  int warp_off = 0;
  for (int i = 0; i < warp; ++i) warp_off += s_w[i];
  int64_t pos = (int64_t)p.block_offsets[(int64_t)b * p.nblk + blockIdx.x] + warp_off + incl - c;
#pragma unroll
  for (int j = 0; j < PP_PER_THREAD; ++j) {
    if (!occ[j]) continue;
    if (pos < p.cap) {
      const float* qp = p.queries + ((int64_t)b * p.Q + q0 + j) * 3;
      float x = __fadd_rn(__fmul_rn(qp[0], p.scale[0]), p.offset[0]);
      float y = __fadd_rn(__fmul_rn(qp[1], p.scale[1]), p.offset[1]);
      float z = __fadd_rn(__fmul_rn(qp[2], p.scale[2]), p.offset[2]);
      if (p.polar2cart) {
        // (r, azimuth deg, elevation deg) -> x = r cos(el) cos(-az), y = r cos(el) sin(-az), z = r sin(el)
        const float d2r = 0.017453292519943295f;
        const float az = -(y * d2r), el = z * d2r, r = x;
        float sa, ca, se, ce;
        sincosf(az, &sa, &ca);
        sincosf(el, &se, &ce);
        x = r * ce * ca;
        y = r * ce * sa;
        z = r * se;
      }
      float* dst = p.points + ((int64_t)b * p.cap + pos) * 3;
      dst[0] = x; dst[1] = y; dst[2] = z;
      if (p.index) p.index[(int64_t)b * p.cap + pos] = (int32_t)(q0 + j);
    }
    ++pos;
  }
}

int occupancy_compact(const float* logits, const float* queries, int B, int64_t Q, float threshold,
                      const float* scale_offset_host, int polar2cart, int64_t cap, float* points, int32_t* index,
                      int32_t* counts, int32_t* block_ws, cudaStream_t stream) {
  RALD_REQUIRE(B > 0 && Q > 0 && Q < (1ll << 31), "occupancy_compact: bad sizes B=%d Q=%lld", B, (long long)Q);
  RALD_REQUIRE(cap > 0, "occupancy_compact: capacity %lld", (long long)cap);
  const int nblk = (int)((Q + PP_BLOCK - 1) / PP_BLOCK);
  dim3 grid((unsigned)nblk, (unsigned)B);
  ProfScope prof(FAM_OTHER, stream, (double)B * Q * 8.0);
  occ_count_kernel<<<grid, PP_THREADS, 0, stream>>>(logits, Q, threshold, nblk, block_ws);
  RALD_LAUNCHED();
  occ_scan_kernel<<<B, 1024, 0, stream>>>(block_ws, nblk, counts);
  RALD_LAUNCHED();
  OccParams p;
  p.logits = logits; p.queries = queries; p.block_offsets = block_ws; p.points = points; p.index = index;
  p.Q = Q; p.cap = cap; p.nblk = nblk; p.thr = threshold; p.polar2cart = polar2cart;
  for (int i = 0; i < 3; ++i) {
    p.scale[i] = scale_offset_host ? scale_offset_host[i] : 1.0f;
    p.offset[i] = scale_offset_host ? scale_offset_host[3 + i] : 0.0f;
  }
  occ_scatter_kernel<<<grid, PP_THREADS, 0, stream>>>(p);
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

extern "C" int rald_occupancy_compact(const float* logits, const float* queries, int B, int64_t Q, float threshold,
                                      const float* scale_offset_host, int polar2cart, int64_t cap, float* points,
                                      int32_t* index, int32_t* counts, int32_t* block_ws, void* stream) {
  return rald::occupancy_compact(logits, queries, B, Q, threshold, scale_offset_host, polar2cart, cap, points, index,
                                 counts, block_ws, static_cast<cudaStream_t>(stream));
}

extern "C" int64_t rald_occupancy_ws_elems(int B, int64_t Q) {
  return (int64_t)B * ((Q + rald::PP_BLOCK - 1) / rald::PP_BLOCK);
}
