// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, M=128, SS operands resident in shared memory, no TMA,
// no epilogue) for N in {32,64,128,256}, one or two accumulators. Prints cycles per MMA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rald_b200/csrc tools/micro/mma_rate.cu -o tools/micro/_bin/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace rald;

template <int N, int NACC, bool TS>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(FMT_BF16, 128, N, 0, 0);
    const uint64_t a = make_sdesc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t b = make_sdesc_sw128(smem_u32(smem + 16384), 16, 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t d = tm + ((it * 4 + kk) % NACC) * 256;
        if (TS) mma_f16_ts(d, tm + 448 + 8 * kk, b + 2 * kk, idesc, 1u);
        else mma_f16_ss(d, a + 2 * kk, b + 2 * kk, idesc, 1u);
      }
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int NACC, bool TS>
void run(const char* name, long long* d) {
  auto kern = k<N, NACC, TS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  kern<<<148, 128, 200 * 1024>>>(d, iters);
  kern<<<148, 128, 200 * 1024>>>(d, iters);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double cyc = (double)h / (iters * 4);
  printf("%-28s %7.1f cycles / MMA  -> %6.0f flop/cycle/SM (%s)\n", name, cyc, 2.0 * 128 * N * 16 / cyc,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  run<256, 1, false>("SS N=256 1 acc", d);
  run<256, 2, false>("SS N=256 2 acc", d);
  run<128, 1, false>("SS N=128 1 acc", d);
  run<128, 2, false>("SS N=128 2 acc", d);
  run<64, 2, false>("SS N=64 2 acc", d);
  run<64, 1, false>("SS N=64 1 acc", d);
  run<32, 1, false>("SS N=32 1 acc", d);
  run<32, 2, false>("SS N=32 2 acc", d);
  run<128, 1, true>("TS N=128 1 acc", d);
  run<128, 2, true>("TS N=128 2 acc", d);
  run<256, 1, true>("TS N=256 1 acc", d);
  return 0;
}
