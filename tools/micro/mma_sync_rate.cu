// Microbenchmark: issue rate of the LEGACY warp-level tensor path on sm_100a (mma.sync), which the evaluation-boundary
// kernel (csrc/dit_misc.cu) uses for its 512 <-> 32 projections: clocks per instruction per SM for
//   tf32  mma.sync.m16n8k8.f32.tf32.tf32.f32    (2 048 flop)
//   bf16  mma.sync.m16n8k16.f32.bf16.bf16.f32   (4 096 flop)
// with 4 / 8 / 16 warps per SM and 1 / 2 / 4 / 8 independent accumulators per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/micro/mma_sync_rate.cu -o tools/micro/_bin/mma_sync_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int NACC>
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink, int iters) {
  float d[NACC][4];
#pragma unroll
  for (int j = 0; j < NACC; ++j) d[j][0] = d[j][1] = d[j][2] = d[j][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  uint32_t b[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NACC; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  if (s == 1.2345f) sink[0] = s;
}

template <int KIND, int NACC>
static void run(const char* name, int warps) {
  long long* out; float* sink;
  cudaMalloc(&out, 8 * 148); cudaMalloc(&sink, 4);
  const int iters = 4096;
  k<KIND, NACC><<<148, warps * 32>>>(out, sink, 16);
  k<KIND, NACC><<<148, warps * 32>>>(out, sink, iters);
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double clk = 0; for (int i = 0; i < 148; ++i) clk += h[i]; clk /= 148;
  const double per_sm = clk / ((double)iters * NACC * warps);
  printf("%-5s warps %2d  acc %d : %7.2f clk / MMA / SM   (%6.1f flop/clk/SM)  err=%s\n", name, warps, NACC, per_sm,
         (KIND == 0 ? 2048.0 : 4096.0) / per_sm, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0, 1>("tf32", w); run<0, 2>("tf32", w); run<0, 4>("tf32", w); run<0, 8>("tf32", w);
    run<1, 1>("bf16", w); run<1, 2>("bf16", w); run<1, 4>("bf16", w); run<1, 8>("bf16", w);
  }
  return 0;
}
