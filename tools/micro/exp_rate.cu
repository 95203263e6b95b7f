// Microbenchmark: per-SM throughput of the exponential recipes a softmax over TMEM scores can use (elements per clock
// per SM, 8 warps = 2 per scheduler as in attn_d64_kernel's softmax groups, and 16 warps):
//   f32     ex2.approx.ftz.f32 per element
//   f16x2   two FFMA (x = s*c - m), cvt.rn.f16x2.f32, ONE ex2.approx.f16x2 per pair (the kernel's recipe)
//   cvt     cvt.rn.f16x2.f32 alone          ex2h    ex2.approx.f16x2 alone
//   poly    Cody-Waite + degree-3 polynomial on the FMA pipe (no MUFU), result packed to f16x2
//   mix     half the pairs through f16x2 MUFU, half through the polynomial
//   f32+cvt two FFMA, two fp32 MUFU exp2, one cvt.rn.f16x2.f32 pack per pair
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rald_b200/csrc tools/micro/exp_rate.cu -o tools/micro/_bin/exp_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace rald;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h_v(uint32_t x) { uint32_t r; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x)); return r; }
__device__ __forceinline__ uint32_t cvt_v(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }

// 2^x for x <= 0 on the FMA pipe: n = round(x), f = x - n in [-0.5, 0.5], p(f) ~ 2^f (degree 3), result = p * 2^n
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                 // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float n = t - 12582912.0f;
  const float f = x - n;
  float p = fmaf(f, 0.05550410866f, 0.2402265070f);
  p = fmaf(p, f, 0.6931471806f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink, int iters, float c, float m) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = -0.01f * (threadIdx.x + j);
  uint32_t acc = 0;
  float facc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float a = v[2 * j], b = v[2 * j + 1];
      if (MODE == 0) { facc += ex2f(fmaf(a, c, -m)) + ex2f(fmaf(b, c, -m)); }
      else if (MODE == 1) { acc ^= ex2h_v(cvt_v(fmaf(a, c, -m), fmaf(b, c, -m))); }
      else if (MODE == 2) { acc ^= cvt_v(a, b); }
      else if (MODE == 3) { acc ^= ex2h_v(__float_as_uint(a) + j); }
      else if (MODE == 4) { acc ^= cvt_v(exp2_poly(fmaf(a, c, -m)), exp2_poly(fmaf(b, c, -m))); }
      else if (MODE == 6) { acc ^= cvt_v(ex2f(fmaf(a, c, -m)), ex2f(fmaf(b, c, -m))); }     // attn_d64's round-2 recipe
      else if (MODE == 5) {
        if (j & 1) acc ^= cvt_v(exp2_poly(fmaf(a, c, -m)), exp2_poly(fmaf(b, c, -m)));
        else acc ^= ex2h_v(cvt_v(fmaf(a, c, -m), fmaf(b, c, -m)));
      }
      v[2 * j] = a + __uint_as_float(acc & 1u) * 1e-30f + facc * 1e-30f;   // loop-carried: nothing is hoisted
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345u || facc == 1.2345f) sink[0] = facc + acc;
}

template <int MODE>
void run(const char* name, int threads) {
  long long* d; float* s;
  cudaMalloc(&d, 8); cudaMalloc(&s, 4);
  const int iters = 2000;
  k<MODE><<<1, threads>>>(d, s, 10, 1.3f, 0.5f);
  k<MODE><<<1, threads>>>(d, s, iters, 1.3f, 0.5f);
  long long h;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  const double elems = 32.0 * threads * iters;
  printf("%-10s %3d threads: %7.3f elements/clk/SM  (%6.2f clk per 32-lane pair instruction per scheduler) %s\n", name, threads,
         elems / (double)h, (double)h / iters / 16.0 / (threads / 128.0), cudaGetErrorString(e));
  cudaFree(d); cudaFree(s);
}

int main() {
  for (int threads : {256, 512}) {
    if (threads == 256) {
      run<0>("f32", 256); run<1>("f16x2", 256); run<2>("cvt", 256); run<3>("ex2h", 256); run<4>("poly", 256); run<5>("mix", 256); run<6>("f32+cvt", 256);
    } else {
      run<0>("f32", 512); run<1>("f16x2", 512); run<2>("cvt", 512); run<3>("ex2h", 512); run<4>("poly", 512); run<5>("mix", 512); run<6>("f32+cvt", 512);
    }
  }
  return 0;
}
