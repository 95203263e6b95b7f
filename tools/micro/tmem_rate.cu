// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM. NW warps (4 or 8; warp w reads TMEM lanes
// 32 (w % 4) .. +31) sweep all 512 columns with .32x32b.x32 loads (32 columns x 32 lanes x 4 B = 4 KB per instruction),
// waiting after every load (DEPTH = 1) or after every DEPTH loads. Prints bytes per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rald_b200/csrc tools/micro/tmem_rate.cu -o tools/micro/_bin/tmem_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace rald;

template <int NW, int DEPTH, bool STORE>
__global__ void __launch_bounds__(NW * 32, 1) k(long long* out, unsigned* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t col0 = (NW == 8 && warp >= 4) ? 256u : 0u;     // 8 warps: each pair member sweeps one column half
  const int ncol = NW == 8 ? 256 : 512;
  unsigned acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < ncol; c += 32 * DEPTH) {
      uint32_t v[DEPTH][32];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) {
        if (STORE) { uint32_t s[16]; for (int j = 0; j < 16; ++j) s[j] = acc + j; tmem_st16(tm + col0 + c + 32 * d, s); }
        else tmem_ld32(tm + col0 + c + 32 * d, v[d]);
      }
      if (STORE) tmem_st_wait(); else tmem_ld_wait();
      if (!STORE) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) acc ^= v[d][0] ^ v[d][31];
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

template <int NW, int DEPTH, bool STORE>
void run(const char* name, int grid) {
  long long* d; unsigned* s;
  cudaMalloc(&d, grid * sizeof(long long)); cudaMalloc(&s, 4);
  const int iters = 200;
  k<NW, DEPTH, STORE><<<grid, NW * 32>>>(d, s, 2);
  k<NW, DEPTH, STORE><<<grid, NW * 32>>>(d, s, iters);
  long long h[148];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  // bytes moved per CTA: 128 lanes x 512 columns x 4 B per sweep (stores: x16 per instruction -> half the columns)
  const double bytes = (STORE ? 0.5 : 1.0) * 128.0 * 512 * 4 * iters;
  printf("%-44s grid %3d: %8.1f cycles per 256 KB sweep, %6.1f B/clk/SM (%s)\n", name, grid, (double)h[0] / iters / (STORE ? 0.5 : 1.0),
         bytes / (double)h[0], cudaGetErrorString(e));
  cudaFree(d); cudaFree(s);
}

int main() {
  run<4, 1, false>("ld x32, 4 warps, wait per load", 1);
  run<4, 2, false>("ld x32, 4 warps, wait per 2 loads", 1);
  run<4, 4, false>("ld x32, 4 warps, wait per 4 loads", 1);
  run<8, 1, false>("ld x32, 8 warps, wait per load", 1);
  run<8, 2, false>("ld x32, 8 warps, wait per 2 loads", 1);
  run<8, 4, false>("ld x32, 8 warps, wait per 4 loads", 1);
  run<8, 4, false>("ld x32, 8 warps, wait per 4 loads", 148);
  run<4, 4, true>("st x16, 4 warps, wait per 4 stores", 1);
  run<8, 4, true>("st x16, 8 warps, wait per 4 stores", 1);
  return 0;
}
