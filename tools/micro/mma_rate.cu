// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, M=128, SS operands resident in shared memory, no TMA,
// no epilogue) for N in {32,64,128,256}, one or two accumulators. Prints cycles per MMA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rald_b200/csrc tools/micro/mma_rate.cu -o tools/micro/_bin/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace rald;

// MODE 0: MMAs only (one commit at the end). MODE 1: a tcgen05.commit to a scratch mbarrier after every 4 MMAs (what a
// GEMM main loop does per 64-wide k-block to free its ring stage). MODE 2: additionally a try_wait on an already
// completed mbarrier + tcgen05.fence::after_thread_sync before every 4 MMAs (the "operands landed" wait).
template <int N, int NACC, bool TS, int MODE = 0>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t scratch_bar[8];
  __shared__ uint64_t ready_bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&scratch_bar[i], 1);
    mbar_init(&ready_bar, 1);
    fence_barrier_init();
    mbar_arrive(&ready_bar);   // phase 0 complete: waits with parity 0 succeed immediately
  }
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
      if (MODE == 2) {
        mbar_wait(&ready_bar, 0);
        tc_fence_after();
      } else if (MODE == 3) {          // wait only
        mbar_wait(&ready_bar, 0);
      } else if (MODE == 4) {          // fence only
        tc_fence_after();
      } else if (MODE == 5) {          // test_wait (non-blocking probe) + fence
        uint32_t ok;
        do {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok) : "r"(smem_u32(&ready_bar)), "r"(0u) : "memory");
        } while (!ok);
        tc_fence_after();
      } else if (MODE == 6) {          // relaxed try_wait, no fence
        uint32_t ok;
        do {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok) : "r"(smem_u32(&ready_bar)), "r"(0u) : "memory");
        } while (!ok);
      }
      if (MODE == 7) {
        // probe of the NEXT wait issued before this group's MMAs; its predicate is consumed after the commit
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred pw, pa;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 pw, [%1], %2;\n\t"
            "setp.ne.b32 pa, %3, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%4], %5, %6, %7, pa;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%4], %8, %9, %7, pa;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%4], %10, %11, %7, pa;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%4], %12, %13, %7, pa;\n\t"
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%14];\n\t"
            "selp.u32 %0, 1, 0, pw;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(&ready_bar)), "r"(0u), "r"(1u), "r"(tm), "l"(a), "l"(b), "r"(idesc), "l"(a + 2), "l"(b + 2),
              "l"(a + 4), "l"(b + 4), "l"(a + 6), "l"(b + 6), "r"(smem_u32(&scratch_bar[it & 7]))
            : "memory");
        if (!ok) mbar_wait(&ready_bar, 0);
        continue;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint32_t d = tm + ((it * 4 + kk) % NACC) * 256;
        if (TS) mma_f16_ts(d, tm + 448 + 8 * kk, b + 2 * kk, idesc, 1u);
        else mma_f16_ss(d, a + 2 * kk, b + 2 * kk, idesc, 1u);
      }
      if (MODE >= 1) tc_commit(&scratch_bar[it & 7]);
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

template <int N, int NACC, bool TS, int MODE = 0>
void run(const char* name, long long* d) {
  auto kern = k<N, NACC, TS, MODE>;
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
  run<32, 1, false, 1>("SS N=32 commit/4", d);
  run<128, 1, false, 1>("SS N=128 commit/4", d);
  run<256, 1, false, 1>("SS N=256 commit/4", d);
  run<32, 1, false, 2>("SS N=32 wait+commit/4", d);
  run<128, 1, false, 2>("SS N=128 wait+commit/4", d);
  run<256, 1, false, 2>("SS N=256 wait+commit/4", d);
  run<32, 1, false, 3>("SS N=32 wait only /4", d);
  run<32, 1, false, 4>("SS N=32 fence only /4", d);
  run<32, 1, false, 5>("SS N=32 test_wait+fence /4", d);
  run<32, 1, false, 6>("SS N=32 relaxed try_wait /4", d);
  run<32, 1, false, 7>("SS N=32 pipelined wait /4", d);
  run<128, 1, false, 7>("SS N=128 pipelined wait /4", d);
  run<128, 1, true>("TS N=128 1 acc", d);
  run<128, 2, true>("TS N=128 2 acc", d);
  run<256, 1, true>("TS N=256 1 acc", d);
  return 0;
}
