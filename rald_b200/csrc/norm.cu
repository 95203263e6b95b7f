// Row LayerNorm kernels (HBM-bound, one warp per 512-wide row, 128-bit loads, fp32 statistics):
//   ln_rows: out_bf16 = LN(x) * g + b with
//     - adaLN modulation  g = 1 + scale[f], b = shift[f]   (model/models_radar_generation.py:127-131), or
//     - affine LayerNorm  g = weight,      b = bias        (model/models_ae.py:38-47, nn.LayerNorm eps 1e-5).
// Input is the fp32 residual stream; output is the bf16 A operand of the following tcgen05 GEMM, so the
// pass moves 4 + 2 bytes per element.
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

template <int D, bool OUT_F32>
__global__ void __launch_bounds__(256)
ln_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
               const float* __restrict__ beta, int64_t mod_frame_stride, int rows_per_frame, int gamma_plus_one,
               void* __restrict__ out, int64_t ldo, int64_t rows, float eps) {
  constexpr int NCH = D / 128;  // float4 chunks per lane
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t warps_total = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  // The modulation / affine vectors are constants of the step (adaLN table: written at least two launches earlier;
  // a kernel is only launched once its predecessor is past its own dependency wait, i.e. once everything before THAT
  // has completed): fetch them for this warp's first row BEFORE the dependency wait. At small batch a warp normalises one
  // row, and this L2 round trip (after the statistics) was a serial ~0.5 us of each of the 72 LayerNorms of an evaluation.
  const float one = gamma_plus_one ? 1.0f : 0.0f;
  float4 g[NCH], b[NCH];
  int64_t f_held = -1;
  if (warp_global < rows) {
    f_held = rows_per_frame > 0 ? warp_global / rows_per_frame : 0;
    const float4* g4 = reinterpret_cast<const float4*>(gamma + f_held * mod_frame_stride);
    const float4* b4 = reinterpret_cast<const float4*>(beta + f_held * mod_frame_stride);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      g[j] = __ldg(g4 + j * 32 + lane);
      b[j] = __ldg(b4 + j * 32 + lane);
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  for (int64_t row = warp_global; row < rows; row += warps_total) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) v[j] = xr[j * 32 + lane];
    const int64_t f = rows_per_frame > 0 ? row / rows_per_frame : 0;
    if (f != f_held) {   // warp-uniform
      f_held = f;
      const float4* g4 = reinterpret_cast<const float4*>(gamma + f * mod_frame_stride);
      const float4* b4 = reinterpret_cast<const float4*>(beta + f * mod_frame_stride);
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        g[j] = __ldg(g4 + j * 32 + lane);
        b[j] = __ldg(b4 + j * 32 + lane);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
      const float o0 = v[j].x * rstd * (g[j].x + one) + b[j].x;
      const float o1 = v[j].y * rstd * (g[j].y + one) + b[j].y;
      const float o2 = v[j].z * rstd * (g[j].z + one) + b[j].z;
      const float o3 = v[j].w * rstd * (g[j].w + one) + b[j].w;
      if (OUT_F32) {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * ldo)[j * 32 + lane] =
            make_float4(o0, o1, o2, o3);
      } else {
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * ldo)[j * 32 + lane] =
            make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
      }
    }
  }
}

int ln_rows(const float* x, int64_t ldx, const float* gamma, const float* beta, int64_t mod_frame_stride,
            int rows_per_frame, int gamma_plus_one, void* out, int64_t ldo, int out_f32, int64_t rows, int D,
            float eps, cudaStream_t stream) {
  RALD_REQUIRE(D == 512, "ln_rows: D=%d unsupported (512 only)", D);
  RALD_REQUIRE(rows > 0, "ln_rows: no rows");
  RALD_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0 && mod_frame_stride % 4 == 0, "ln_rows: strides must be 16-byte multiples");
  const int warps_per_block = 8;
  int64_t blocks = (rows + warps_per_block - 1) / warps_per_block;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  ProfScope prof(FAM_LN, stream, (double)rows * D * (out_f32 ? 8.0 : 6.0));
  if (out_f32)
    RALD_CHECK_CUDA(launch_pdl(ln_rows_kernel<512, true>, dim3((unsigned)blocks), dim3(256), 0, stream, x, ldx, gamma, beta,
                               mod_frame_stride, rows_per_frame, gamma_plus_one, out, ldo, rows, eps));
  else
    RALD_CHECK_CUDA(launch_pdl(ln_rows_kernel<512, false>, dim3((unsigned)blocks), dim3(256), 0, stream, x, ldx, gamma,
                               beta, mod_frame_stride, rows_per_frame, gamma_plus_one, out, ldo, rows, eps));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald
