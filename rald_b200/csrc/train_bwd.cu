// Backward-pass helpers of the denoiser's training step (EDMLoss under autograd,
// model/models_radar_generation.py:277-295, engine_generation.py:89-110). The matrix products of the backward pass run
// on the same tcgen05 GEMM as the forward pass (gemm.cu): dgrad = dY W with a transposed bf16 copy of the weight as the
// K-major B operand, wgrad = dY^T X as a GEMM over K = rows with both activations transposed and the fp32 gradient
// accumulated by the TMA reduce-add epilogue. This file holds what surrounds them:
//   cast_transpose_kernel   fp32 / bf16 [R, C] -> bf16 [R, C] and / or bf16 [C, R]   (operands of the wgrad GEMMs)
//   colsum_*                bias gradients: column sums over rows, two deterministic stages
//   ln_bwd_kernel           LayerNorm / adaLN backward (:119-131): dx into the residual-stream gradient, per-64-row
//                           partial sums of d(scale|weight), d(shift|bias), reduced per frame by ln_bwd_reduce_kernel
//   geglu_fwd / geglu_bwd   GEGLU (:88-95) on a materialised projection (the training forward keeps u = xW^T + b)
//   sgemm_f32_kernel        small fp32 GEMM with optional transposes (timestep-embedding MLP, [B, 512] operands)
//   radar_tokens_bwd_kernel gradients of the token projection and the r / a / e embeddings (:390-405)
#include "../../include/rald_b200.h"

#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace rald {

__device__ __forceinline__ float bf16_bits_to_f32(uint16_t v) { return __uint_as_float(static_cast<uint32_t>(v) << 16); }
__device__ __forceinline__ uint16_t f32_to_bf16_bits(float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// ---------------------------------------------------------------------------------------------------
// in [R, C] (fp32 or bf16, pitch ld_in) -> out [R, C] bf16 (optional) and out_t [C, R] bf16 (optional, pitch ld_t).
// 64 x 64 tiles through shared memory, both global sides coalesced.
// ---------------------------------------------------------------------------------------------------
template <bool IN_F32>
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const void* __restrict__ in, int64_t ld_in, int64_t R, int64_t C, uint16_t* __restrict__ out,
                      int64_t ld_out, uint16_t* __restrict__ out_t, int64_t ld_t) {
  __shared__ uint16_t tile[64][66];
  const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;   // 64 x 4
#pragma unroll 4
  for (int i = ty; i < 64; i += 4) {
    const int64_t r = r0 + i, c = c0 + tx;
    uint16_t v = 0;
    if (r < R && c < C) {
      if (IN_F32) v = f32_to_bf16_bits(reinterpret_cast<const float*>(in)[r * ld_in + c]);
      else v = reinterpret_cast<const uint16_t*>(in)[r * ld_in + c];
      if (out != nullptr) out[r * ld_out + c] = v;
    }
    tile[i][tx] = v;
  }
  if (out_t == nullptr) return;
  __syncthreads();
#pragma unroll 4
  for (int i = ty; i < 64; i += 4) {
    const int64_t c = c0 + i, r = r0 + tx;
    if (c < C && r < R) out_t[c * ld_t + r] = tile[tx][i];
  }
}

// Fast path for full, aligned 64 x 64 tiles: 128-bit global accesses on both sides (8 elements per thread per access)
// and, optionally, the per-tile column sums of the fp32 INPUT (bias gradients: the same pass that produces the wgrad
// operands, instead of another sweep over dY): colpart[blockIdx.y][c] = sum of the tile's 64 rows of column c.
template <bool IN_F32>
__global__ void __launch_bounds__(256)
cast_transpose_vec_kernel(const void* __restrict__ in, int64_t ld_in, int64_t R, int64_t C, uint16_t* __restrict__ out,
                          int64_t ld_out, uint16_t* __restrict__ out_t, int64_t ld_t, float* __restrict__ colpart) {
  __shared__ uint16_t tile[64][72];          // 144-byte rows: 16-byte aligned vector stores
  __shared__ float csum[32][64];
  const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
  const int cv = threadIdx.x & 7, rr = threadIdx.x >> 3;     // 8 column octets x 32 rows, two passes
  float cs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) cs[k] = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = rr + 32 * i;
    uint4 packed;
    if (IN_F32) {
      const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + (r0 + row) * ld_in + c0 + cv * 8);
      const float4 a = src[0], b = src[1];
      cs[0] += a.x; cs[1] += a.y; cs[2] += a.z; cs[3] += a.w;
      cs[4] += b.x; cs[5] += b.y; cs[6] += b.z; cs[7] += b.w;
      packed = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    } else {
      packed = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(in) + (r0 + row) * ld_in + c0 + cv * 8);
      if (colpart != nullptr) {
        const uint32_t u[4] = {packed.x, packed.y, packed.z, packed.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          cs[2 * k] += __uint_as_float(u[k] << 16);
          cs[2 * k + 1] += __uint_as_float(u[k] & 0xffff0000u);
        }
      }
    }
    if (out != nullptr) *reinterpret_cast<uint4*>(out + (r0 + row) * ld_out + c0 + cv * 8) = packed;
    *reinterpret_cast<uint4*>(&tile[row][cv * 8]) = packed;
  }
  if (colpart != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) csum[rr][cv * 8 + k] = cs[k];
  }
  __syncthreads();
  if (colpart != nullptr && threadIdx.x < 64) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) a += csum[k][threadIdx.x];
    colpart[(int64_t)blockIdx.y * C + c0 + threadIdx.x] = a;
  }
  if (out_t == nullptr) return;
  // transposed side: 8 threads write the 128 contiguous bytes (64 rows) of one output row = input column
  const int rv = threadIdx.x & 7, cc = threadIdx.x >> 3;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int col = cc + 32 * i;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      w[k] = (uint32_t)tile[rv * 8 + 2 * k][col] | ((uint32_t)tile[rv * 8 + 2 * k + 1][col] << 16);
    *reinterpret_cast<uint4*>(out_t + (c0 + col) * ld_t + r0 + rv * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

int cast_transpose(const void* in, int in_f32, int64_t ld_in, int64_t R, int64_t C, void* out, int64_t ld_out,
                   void* out_t, int64_t ld_t, float* colpart, cudaStream_t stream) {
  RALD_REQUIRE(R > 0 && C > 0, "cast_transpose: bad shape");
  RALD_REQUIRE(out != nullptr || out_t != nullptr || colpart != nullptr, "cast_transpose: no output");
  dim3 grid((unsigned)((C + 63) / 64), (unsigned)((R + 63) / 64));
  RALD_REQUIRE(grid.y < 65536, "cast_transpose: too many rows (%lld)", (long long)R);
  const int esz = in_f32 ? 4 : 2;
  const bool vec = R % 64 == 0 && C % 64 == 0 && (ld_in * esz) % 16 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                   (out == nullptr || (ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)) &&
                   (out_t == nullptr || (ld_t % 8 == 0 && (reinterpret_cast<uintptr_t>(out_t) & 15) == 0));
  RALD_REQUIRE(colpart == nullptr || vec, "cast_transpose: column sums need the aligned 64 x 64-tile form");
  uint16_t* o = reinterpret_cast<uint16_t*>(out);
  uint16_t* ot = reinterpret_cast<uint16_t*>(out_t);
  if (vec) {
    if (in_f32) cast_transpose_vec_kernel<true><<<grid, 256, 0, stream>>>(in, ld_in, R, C, o, ld_out, ot, ld_t, colpart);
    else cast_transpose_vec_kernel<false><<<grid, 256, 0, stream>>>(in, ld_in, R, C, o, ld_out, ot, ld_t, colpart);
  } else {
    if (in_f32) cast_transpose_kernel<true><<<grid, 256, 0, stream>>>(in, ld_in, R, C, o, ld_out, ot, ld_t);
    else cast_transpose_kernel<false><<<grid, 256, 0, stream>>>(in, ld_in, R, C, o, ld_out, ot, ld_t);
  }
  RALD_LAUNCHED();
  return 0;
}

// out[c] (+)= sum_k partial[k][c] (fixed order): second stage of the column sums produced by cast_transpose
int colsum_finish(const float* partial, int chunks, int64_t C, float* out, int accumulate, cudaStream_t stream);

// out[f][j][c] = bf16(in[f][j][c] - mean_j in[f][j][c]) for fp16 in: the V columns the forward attention consumed as
// fp16, centred per (frame, column) and re-encoded for the backward attention kernel (attn_bwd.cu explains why).
// One CTA per (frame, 64 columns): 4 row lanes x 64 columns, mean in fp32 with a fixed summation order.
__global__ void __launch_bounds__(256)
center_cast_f16_bf16_kernel(const __half* __restrict__ in, int64_t ld_in, uint16_t* __restrict__ out, int64_t ld_out,
                            int rows_per_frame, int64_t C) {
  __shared__ float part[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int64_t c = (int64_t)blockIdx.x * 64 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_frame;
  float a = 0.f;
  if (c < C)
    for (int j = ty; j < rows_per_frame; j += 4) a += __half2float(in[(r0 + j) * ld_in + c]);
  part[ty][tx] = a;
  __syncthreads();
  const float mean = ((part[0][tx] + part[1][tx]) + (part[2][tx] + part[3][tx])) / (float)rows_per_frame;
  if (c < C)
    for (int j = ty; j < rows_per_frame; j += 4)
      out[(r0 + j) * ld_out + c] = f32_to_bf16_bits(__half2float(in[(r0 + j) * ld_in + c]) - mean);
}

int center_cast_f16_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int frames, int rows_per_frame,
                         int64_t C, cudaStream_t stream) {
  RALD_REQUIRE(frames > 0 && rows_per_frame > 0 && C > 0 && frames < 65536, "center_cast_f16_bf16: bad shape");
  dim3 grid((unsigned)((C + 63) / 64), (unsigned)frames);
  center_cast_f16_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __half*>(in), ld_in,
                                                        reinterpret_cast<uint16_t*>(out), ld_out, rows_per_frame, C);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// column sums: partial[chunk][c] = sum of rows [chunk*rows_per_chunk, ...) ; out[c] (+)= sum_chunk partial (fixed order)
// ---------------------------------------------------------------------------------------------------
template <bool IN_F32>
__global__ void __launch_bounds__(128)
colsum_partial_kernel(const void* __restrict__ in, int64_t ld, int64_t R, int64_t C, int64_t rows_per_chunk,
                      float* __restrict__ partial) {
  const int64_t c = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r1 = min(r0 + rows_per_chunk, R);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int64_t r = r0;
  auto ld1 = [&](int64_t rr) -> float {
    if (IN_F32) return reinterpret_cast<const float*>(in)[rr * ld + c];
    return bf16_bits_to_f32(reinterpret_cast<const uint16_t*>(in)[rr * ld + c]);
  };
  for (; r + 4 <= r1; r += 4) {
    a0 += ld1(r); a1 += ld1(r + 1); a2 += ld1(r + 2); a3 += ld1(r + 3);
  }
  for (; r < r1; ++r) a0 += ld1(r);
  partial[(int64_t)blockIdx.y * C + c] = (a0 + a1) + (a2 + a3);
}

__global__ void __launch_bounds__(128)
colsum_final_kernel(const float* __restrict__ partial, int chunks, int64_t C, float* __restrict__ out, int accumulate) {
  const int64_t c = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  float a = 0.f;
  for (int k = 0; k < chunks; ++k) a += partial[(int64_t)k * C + c];
  out[c] = accumulate ? out[c] + a : a;
}

int colsum_finish(const float* partial, int chunks, int64_t C, float* out, int accumulate, cudaStream_t stream) {
  RALD_REQUIRE(chunks > 0 && C > 0, "colsum_finish: bad shape");
  colsum_final_kernel<<<(unsigned)((C + 127) / 128), 128, 0, stream>>>(partial, chunks, C, out, accumulate);
  RALD_LAUNCHED();
  return 0;
}

int colsum(const void* in, int in_f32, int64_t ld, int64_t R, int64_t C, float* partial_ws, int64_t ws_elems, float* out,
           int accumulate, cudaStream_t stream) {
  RALD_REQUIRE(R > 0 && C > 0, "colsum: bad shape");
  int64_t chunks = (R + 255) / 256;
  if (chunks > 512) chunks = 512;
  const int64_t rpc = (R + chunks - 1) / chunks;
  chunks = (R + rpc - 1) / rpc;
  RALD_REQUIRE(ws_elems >= chunks * C, "colsum: workspace of %lld floats < %lld", (long long)ws_elems,
               (long long)(chunks * C));
  dim3 grid((unsigned)((C + 127) / 128), (unsigned)chunks);
  if (in_f32) colsum_partial_kernel<true><<<grid, 128, 0, stream>>>(in, ld, R, C, rpc, partial_ws);
  else colsum_partial_kernel<false><<<grid, 128, 0, stream>>>(in, ld, R, C, rpc, partial_ws);
  RALD_LAUNCHED();
  colsum_final_kernel<<<(unsigned)((C + 127) / 128), 128, 0, stream>>>(partial_ws, (int)chunks, C, out, accumulate);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward over 512-wide rows. y = xhat * g + b with g = 1 + scale[f] (adaLN) or g = weight.
//   dxhat = dy * g ; dx = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
//   dh[row] (+)= dx ; partial[cta][0][c] = sum_rows dy * xhat ; partial[cta][1][c] = sum_rows dy     (64 rows per CTA)
// One warp per row (lane owns columns j*128 + lane*4 .. +3, j = 0..3, the layout of ln_rows_kernel); statistics are
// recomputed from x exactly as the forward pass computes them.
// ---------------------------------------------------------------------------------------------------
constexpr int LNB_ROWS = 64;
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ x, const uint16_t* __restrict__ dy, const float* __restrict__ gamma,
              int64_t mod_frame_stride, int rows_per_frame, int gamma_plus_one, float* __restrict__ dh, int accumulate,
              float* __restrict__ partial, int64_t rows, float eps) {
  constexpr int D = 512;
  __shared__ float red[8][2][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = (int64_t)blockIdx.x * LNB_ROWS;
  float pg[16], pb[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { pg[i] = 0.f; pb[i] = 0.f; }
  const float one = gamma_plus_one ? 1.0f : 0.0f;
  for (int i = warp; i < LNB_ROWS; i += 8) {
    const int64_t row = row0 + i;
    if (row >= rows) break;
    const float4* xr = reinterpret_cast<const float4*>(x + row * D);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + row * D);
    float4 v[4];
    float dyv[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = xr[j * 32 + lane];
      const uint2 d = dr[j * 32 + lane];
      dyv[4 * j + 0] = __uint_as_float(d.x << 16);
      dyv[4 * j + 1] = __uint_as_float(d.x & 0xffff0000u);
      dyv[4 * j + 2] = __uint_as_float(d.y << 16);
      dyv[4 * j + 3] = __uint_as_float(d.y & 0xffff0000u);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
    const int64_t f = rows_per_frame > 0 ? row / rows_per_frame : 0;
    const float4* g4 = reinterpret_cast<const float4*>(gamma + f * mod_frame_stride);
    float xh[16], dxh[16];
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 g = __ldg(g4 + j * 32 + lane);
      xh[4 * j + 0] = v[j].x * rstd; xh[4 * j + 1] = v[j].y * rstd;
      xh[4 * j + 2] = v[j].z * rstd; xh[4 * j + 3] = v[j].w * rstd;
      dxh[4 * j + 0] = dyv[4 * j + 0] * (g.x + one); dxh[4 * j + 1] = dyv[4 * j + 1] * (g.y + one);
      dxh[4 * j + 2] = dyv[4 * j + 2] * (g.z + one); dxh[4 * j + 3] = dyv[4 * j + 3] * (g.w + one);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      m1 += dxh[k];
      m2 = fmaf(dxh[k], xh[k], m2);
      pg[k] = fmaf(dyv[k], xh[k], pg[k]);
      pb[k] += dyv[k];
    }
    m1 = warp_sum(m1) * (1.0f / D);
    m2 = warp_sum(m2) * (1.0f / D);
    float4* dhr = reinterpret_cast<float4*>(dh + row * D);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 o;
      o.x = rstd * (dxh[4 * j + 0] - m1 - xh[4 * j + 0] * m2);
      o.y = rstd * (dxh[4 * j + 1] - m1 - xh[4 * j + 1] * m2);
      o.z = rstd * (dxh[4 * j + 2] - m1 - xh[4 * j + 2] * m2);
      o.w = rstd * (dxh[4 * j + 3] - m1 - xh[4 * j + 3] * m2);
      if (accumulate) {
        const float4 p = dhr[j * 32 + lane];
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      dhr[j * 32 + lane] = o;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      red[warp][0][j * 128 + lane * 4 + k] = pg[4 * j + k];
      red[warp][1][j * 128 + lane * 4 + k] = pb[4 * j + k];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += 256) {
    const int which = i / D, c = i % D;
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[w][which][c];
    partial[((int64_t)blockIdx.x * 2 + which) * D + c] = a;
  }
}

// out[g][which][c] (+)= sum over the ctas_per_group partials of group g; out pitch = out_group_stride floats per group,
// `which` stride = which_stride (adaLN table rows hold scale | shift = 512 apart; affine LN: two separate vectors)
__global__ void __launch_bounds__(256)
ln_bwd_reduce_kernel(const float* __restrict__ partial, int ctas_per_group, float* __restrict__ out,
                     int64_t out_group_stride, int64_t which_stride, int accumulate) {
  constexpr int D = 512;
  const int g = blockIdx.x;
  for (int i = threadIdx.x; i < 2 * D; i += 256) {
    const int which = i / D, c = i % D;
    float a = 0.f;
    for (int k = 0; k < ctas_per_group; ++k) a += partial[(((int64_t)g * ctas_per_group + k) * 2 + which) * D + c];
    float* o = out + g * out_group_stride + which * which_stride + c;
    *o = accumulate ? *o + a : a;
  }
}

int ln_bwd(const float* x, const void* dy_bf16, const float* gamma, int64_t mod_frame_stride, int rows_per_frame,
           int gamma_plus_one, float* dh, int accumulate_dh, float* partial_ws, int64_t ws_elems, float* dparam,
           int64_t dparam_group_stride, int64_t dparam_which_stride, int accumulate_dparam, int64_t rows, int D,
           float eps, cudaStream_t stream) {
  RALD_REQUIRE(D == 512, "ln_bwd: D=%d unsupported (512 only)", D);
  RALD_REQUIRE(rows > 0 && rows % LNB_ROWS == 0, "ln_bwd: rows=%lld must be a positive multiple of %d", (long long)rows,
               LNB_ROWS);
  const int64_t group_rows = rows_per_frame > 0 ? rows_per_frame : rows;
  RALD_REQUIRE(group_rows % LNB_ROWS == 0 && rows % group_rows == 0, "ln_bwd: %lld rows per frame must divide %lld rows "
               "and be a multiple of %d", (long long)group_rows, (long long)rows, LNB_ROWS);
  const int64_t ctas = rows / LNB_ROWS;
  RALD_REQUIRE(ws_elems >= ctas * 2 * D, "ln_bwd: workspace of %lld floats < %lld", (long long)ws_elems,
               (long long)(ctas * 2 * D));
  ln_bwd_kernel<<<(unsigned)ctas, 256, 0, stream>>>(x, reinterpret_cast<const uint16_t*>(dy_bf16), gamma, mod_frame_stride,
                                                    rows_per_frame, gamma_plus_one, dh, accumulate_dh, partial_ws, rows,
                                                    eps);
  RALD_LAUNCHED();
  const int groups = (int)(rows / group_rows);
  ln_bwd_reduce_kernel<<<groups, 256, 0, stream>>>(partial_ws, (int)(group_rows / LNB_ROWS), dparam, dparam_group_stride,
                                                   dparam_which_stride, accumulate_dparam);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// GEGLU on a materialised projection u [T, 2*inner] bf16 (value columns [0, inner), gate columns [inner, 2 inner)),
// erf GELU as F.gelu (models_radar_generation.py:94-95).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_exact_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// 8 outputs per thread: 128-bit loads of the value and the gate octet (inner % 8 == 0)
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(const uint4* __restrict__ u, int64_t T, int inner, uint4* __restrict__ g) {
  const int64_t oct = inner / 8;             // octets per half row
  const int64_t n = T * oct;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t t = i / oct, c = i - t * oct;
    const uint4 v = u[t * (2 * oct) + c], gt = u[t * (2 * oct) + oct + c];
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {gt.x, gt.y, gt.z, gt.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = pack_bf16x2(__uint_as_float(vv[k] << 16) * gelu_exact_f(__uint_as_float(gg[k] << 16)),
                         __uint_as_float(vv[k] & 0xffff0000u) * gelu_exact_f(__uint_as_float(gg[k] & 0xffff0000u)));
    g[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const uint4* __restrict__ u, const uint4* __restrict__ dg, int64_t T, int inner, uint4* __restrict__ du) {
  const int64_t oct = inner / 8;
  const int64_t n = T * oct;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int64_t t = i / oct, c = i - t * oct;
    const uint4 v = u[t * (2 * oct) + c], gt = u[t * (2 * oct) + oct + c], d = dg[i];
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {gt.x, gt.y, gt.z, gt.w}, dd[4] = {d.x, d.y, d.z, d.w};
    uint32_t ov[4], og[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float v0 = __uint_as_float(vv[k] << 16), v1 = __uint_as_float(vv[k] & 0xffff0000u);
      const float g0 = __uint_as_float(gg[k] << 16), g1 = __uint_as_float(gg[k] & 0xffff0000u);
      const float d0 = __uint_as_float(dd[k] << 16), d1 = __uint_as_float(dd[k] & 0xffff0000u);
      ov[k] = pack_bf16x2(d0 * gelu_exact_f(g0), d1 * gelu_exact_f(g1));
      og[k] = pack_bf16x2(d0 * v0 * gelu_grad_f(g0), d1 * v1 * gelu_grad_f(g1));
    }
    du[t * (2 * oct) + c] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
    du[t * (2 * oct) + oct + c] = make_uint4(og[0], og[1], og[2], og[3]);
  }
}

int geglu_fwd(const void* u, int64_t T, int inner, void* g, cudaStream_t stream) {
  RALD_REQUIRE(T > 0 && inner > 0 && inner % 8 == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(g) & 15) == 0, "geglu_fwd: bad shape / alignment (inner % 8, 16-byte pointers)");
  const int64_t n = T * (inner / 8);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  geglu_fwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(u), T, inner,
                                                         reinterpret_cast<uint4*>(g));
  RALD_LAUNCHED();
  return 0;
}

int geglu_bwd(const void* u, const void* dg, int64_t T, int inner, void* du, cudaStream_t stream) {
  RALD_REQUIRE(T > 0 && inner > 0 && inner % 8 == 0 && ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(dg) |
                                                         reinterpret_cast<uintptr_t>(du)) & 15) == 0,
               "geglu_bwd: bad shape / alignment (inner % 8, 16-byte pointers)");
  const int64_t n = T * (inner / 8);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  geglu_bwd_kernel<<<(unsigned)blocks, 256, 0, stream>>>(reinterpret_cast<const uint4*>(u),
                                                         reinterpret_cast<const uint4*>(dg), T, inner,
                                                         reinterpret_cast<uint4*>(du));
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// C[M, N] = alpha * op(A) op(B) + beta * C in fp32, row-major, op = identity or transpose. 64 x 64 tiles, 16-deep
// k-steps, 4 x 4 outputs per thread. For the [B, 512]-sized operands of the timestep-embedding MLP only.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgemm_f32_kernel(int ta, int tb, int M, int N, int K, float alpha, const float* __restrict__ A, int64_t lda,
                 const float* __restrict__ B, int64_t ldb, float beta, float* __restrict__ C, int64_t ldc) {
  __shared__ float sA[16][64 + 4], sB[16][64 + 4];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i >> 6, mm = i & 63;     // element (m0 + mm, k0 + kk) of op(A), (k0 + kk, n0 + mm) of op(B)
      const int m = m0 + mm, n = n0 + mm, k = k0 + kk;
      float a = 0.f, b = 0.f;
      if (k < K) {
        if (m < M) a = ta ? A[(int64_t)k * lda + m] : A[(int64_t)m * lda + k];
        if (n < N) b = tb ? B[(int64_t)n * ldb + k] : B[(int64_t)k * ldb + n];
      }
      sA[kk][mm] = a;
      sB[kk][mm] = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sA[kk][ty * 4 + i]; b[i] = sB[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float* c = C + (int64_t)m * ldc + n;
        *c = beta == 0.f ? alpha * acc[i][j] : alpha * acc[i][j] + beta * *c;
      }
    }
}

int sgemm_f32(int ta, int tb, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B, int64_t ldb,
              float beta, float* C, int64_t ldc, cudaStream_t stream) {
  RALD_REQUIRE(M > 0 && N > 0 && K > 0, "sgemm_f32: bad shape");
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
  RALD_REQUIRE(grid.y < 65536, "sgemm_f32: M too large");
  sgemm_f32_kernel<<<grid, 256, 0, stream>>>(ta, tb, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Gradients of process_radar_cond's tail (models_radar_generation.py:390-405): tok[t, j] = w[j, :] . feat[t, :] + b[j]
// + r_emb[r(t), j] + a_emb[a(t), j] + e_emb[e(t), j]. One thread per output feature j walks all tokens in order
// (deterministic); the token count is B * 64 and dim 512, so this is a few microseconds of work.
// ---------------------------------------------------------------------------------------------------
constexpr int RT_MAX_CZ = 32, RT_MAX_R = 16, RT_MAX_A = 8, RT_MAX_E = 4;
__global__ void __launch_bounds__(128)
radar_tokens_bwd_kernel(const float* __restrict__ dtok, const float* __restrict__ feat, int64_t ntok, int nr, int na,
                        int ne, int cz, int dim, float* __restrict__ dw, float* __restrict__ db,
                        float* __restrict__ dr, float* __restrict__ da, float* __restrict__ de) {
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j >= dim) return;
  float aw[RT_MAX_CZ], ar[RT_MAX_R], aa[RT_MAX_A], ae[RT_MAX_E];
#pragma unroll
  for (int c = 0; c < RT_MAX_CZ; ++c) aw[c] = 0.f;
#pragma unroll
  for (int c = 0; c < RT_MAX_R; ++c) ar[c] = 0.f;
#pragma unroll
  for (int c = 0; c < RT_MAX_A; ++c) aa[c] = 0.f;
#pragma unroll
  for (int c = 0; c < RT_MAX_E; ++c) ae[c] = 0.f;
  float ab = 0.f;
  for (int64_t t = 0; t < ntok; ++t) {
    const float g = dtok[t * dim + j];
    const int e = (int)(t % ne), a = (int)((t / ne) % na), r = (int)((t / ((int64_t)ne * na)) % nr);
    ab += g;
#pragma unroll
    for (int c = 0; c < RT_MAX_CZ; ++c)
      if (c < cz) aw[c] = fmaf(g, feat[t * cz + c], aw[c]);
#pragma unroll
    for (int c = 0; c < RT_MAX_R; ++c) ar[c] += (c == r) ? g : 0.f;
#pragma unroll
    for (int c = 0; c < RT_MAX_A; ++c) aa[c] += (c == a) ? g : 0.f;
#pragma unroll
    for (int c = 0; c < RT_MAX_E; ++c) ae[c] += (c == e) ? g : 0.f;
  }
  db[j] = ab;
  for (int c = 0; c < cz; ++c) dw[(int64_t)j * cz + c] = aw[c];
  for (int c = 0; c < nr; ++c) dr[(int64_t)c * dim + j] = ar[c];
  for (int c = 0; c < na; ++c) da[(int64_t)c * dim + j] = aa[c];
  for (int c = 0; c < ne; ++c) de[(int64_t)c * dim + j] = ae[c];
}

int radar_tokens_bwd(const float* dtok, const float* feat, int B, int nr, int na, int ne, int cz, int dim, float* dw,
                     float* db, float* dr, float* da, float* de, cudaStream_t stream) {
  RALD_REQUIRE(cz <= RT_MAX_CZ && nr <= RT_MAX_R && na <= RT_MAX_A && ne <= RT_MAX_E,
               "radar_tokens_bwd: geometry cz=%d r=%d a=%d e=%d exceeds %d / %d / %d / %d", cz, nr, na, ne, RT_MAX_CZ,
               RT_MAX_R, RT_MAX_A, RT_MAX_E);
  const int64_t ntok = (int64_t)B * nr * na * ne;
  RALD_REQUIRE(ntok > 0, "radar_tokens_bwd: no tokens");
  radar_tokens_bwd_kernel<<<(unsigned)((dim + 127) / 128), 128, 0, stream>>>(dtok, feat, ntok, nr, na, ne, cz, dim, dw, db,
                                                                            dr, da, de);
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald

// ---------------------------------------------------------------------------------------------------
// C ABI (declared in include/rald_b200.h)
// ---------------------------------------------------------------------------------------------------
extern "C" {

int rald_cast_transpose(const void* in, int in_f32, int64_t ld_in, int64_t R, int64_t C, void* out_bf16, int64_t ld_out,
                        void* out_t_bf16, int64_t ld_t, float* colsum_partial, void* stream) {
  return rald::cast_transpose(in, in_f32, ld_in, R, C, out_bf16, ld_out, out_t_bf16, ld_t, colsum_partial,
                              static_cast<cudaStream_t>(stream));
}

int rald_colsum_finish(const float* partial, int chunks, int64_t C, float* out, int accumulate, void* stream) {
  return rald::colsum_finish(partial, chunks, C, out, accumulate, static_cast<cudaStream_t>(stream));
}

int rald_center_cast_f16_bf16(const void* in_f16, int64_t ld_in, void* out_bf16, int64_t ld_out, int frames,
                              int rows_per_frame, int64_t C, void* stream) {
  return rald::center_cast_f16_bf16(in_f16, ld_in, out_bf16, ld_out, frames, rows_per_frame, C,
                                    static_cast<cudaStream_t>(stream));
}

int rald_colsum(const void* in, int in_f32, int64_t ld, int64_t R, int64_t C, float* partial_ws, int64_t ws_elems,
                float* out, int accumulate, void* stream) {
  return rald::colsum(in, in_f32, ld, R, C, partial_ws, ws_elems, out, accumulate, static_cast<cudaStream_t>(stream));
}

int rald_ln_bwd(const float* x, const void* dy_bf16, const float* gamma, int64_t mod_frame_stride, int rows_per_frame,
                int gamma_plus_one, float* dh, int accumulate_dh, float* partial_ws, int64_t ws_elems, float* dparam,
                int64_t dparam_group_stride, int64_t dparam_which_stride, int accumulate_dparam, int64_t rows, int D,
                float eps, void* stream) {
  return rald::ln_bwd(x, dy_bf16, gamma, mod_frame_stride, rows_per_frame, gamma_plus_one, dh, accumulate_dh, partial_ws,
                      ws_elems, dparam, dparam_group_stride, dparam_which_stride, accumulate_dparam, rows, D, eps,
                      static_cast<cudaStream_t>(stream));
}

int rald_geglu_fwd(const void* u_bf16, int64_t T, int inner, void* g_bf16, void* stream) {
  return rald::geglu_fwd(u_bf16, T, inner, g_bf16, static_cast<cudaStream_t>(stream));
}

int rald_geglu_bwd(const void* u_bf16, const void* dg_bf16, int64_t T, int inner, void* du_bf16, void* stream) {
  return rald::geglu_bwd(u_bf16, dg_bf16, T, inner, du_bf16, static_cast<cudaStream_t>(stream));
}

int rald_sgemm_f32(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                   int64_t ldb, float beta, float* C, int64_t ldc, void* stream) {
  return rald::sgemm_f32(trans_a, trans_b, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc,
                         static_cast<cudaStream_t>(stream));
}

int rald_radar_tokens_bwd(const float* dtok, const float* feat, int B, int nr, int na, int ne, int cz, int dim, float* dw,
                          float* db, float* dr_emb, float* da_emb, float* de_emb, void* stream) {
  return rald::radar_tokens_bwd(dtok, feat, B, nr, na, ne, cz, dim, dw, db, dr_emb, da_emb, de_emb,
                                static_cast<cudaStream_t>(stream));
}

}  // extern "C"
