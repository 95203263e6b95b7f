// Small fp32 SIMT kernels around the denoiser blocks (latency / HBM bound, no tensor-core work):
//   temb_kernel      PositionalEmbedding + map_layer0/1 + SiLU          models_radar_generation.py:27-33, 217-219
//   adaln_kernel     all depth*3 AdaLayerNorm linears for S sigmas      models_radar_generation.py:128-129
//   boundary_kernel  final LayerNorm + proj_out + EDM preconditioning + Euler/Heun update + next proj_in
//                                                                        :230-232, :422-429, :265-273, :221
//   radar_tokens_kernel  token projection + r/a/e embeddings            :390-405
#include <stdlib.h>

#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------
// t_emb[s, :] = silu(W1 silu(W0 [cos(c f), sin(c f)] + b0) + b1),  c = ln(sigma_s) / 4
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
temb_kernel(const float* __restrict__ sigma, const float* __restrict__ freqs, int half, const float* __restrict__ w0,
            const float* __restrict__ b0, const float* __restrict__ w1, const float* __restrict__ b1, int dim,
            float* __restrict__ t_emb) {
  extern __shared__ float sm[];
  float* emb = sm;             // [2*half]
  float* h0 = sm + 2 * half;   // [dim]
  const int s = blockIdx.x;
  const float c_noise = logf(sigma[s]) / 4.0f;
  for (int j = threadIdx.x; j < half; j += blockDim.x) {
    const float x = c_noise * freqs[j];
    emb[j] = cosf(x);
    emb[half + j] = sinf(x);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int in0 = 2 * half;
  for (int r = warp; r < dim; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < in0; k += 32) acc += w0[(int64_t)r * in0 + k] * emb[k];
    acc = warp_sum(acc);
    if (lane == 0) h0[r] = silu(acc + b0[r]);
  }
  __syncthreads();
  for (int r = warp; r < dim; r += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc += w1[(int64_t)r * dim + k] * h0[k];
    acc = warp_sum(acc);
    if (lane == 0) t_emb[(int64_t)s * dim + r] = silu(acc + b1[r]);
  }
}

// mod[s, r] = ada_w[r, :] . t_emb[s, :] + ada_b[r]   for r < R = depth*3*2*dim ; one warp per row r.
constexpr int ADA_SCHUNK = 16;
__global__ void __launch_bounds__(256)
adaln_kernel(const float* __restrict__ t_emb, int S, const float* __restrict__ ada_w, const float* __restrict__ ada_b,
             int64_t R, float* __restrict__ mod) {
  __shared__ float4 te[ADA_SCHUNK][128];
  const int s0 = blockIdx.y * ADA_SCHUNK;
  const int ns = (S - s0) < ADA_SCHUNK ? (S - s0) : ADA_SCHUNK;
  for (int i = threadIdx.x; i < ns * 128; i += blockDim.x)
    te[i / 128][i % 128] = reinterpret_cast<const float4*>(t_emb + (int64_t)(s0 + i / 128) * 512)[i % 128];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= R) return;
  float4 w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) w[j] = __ldg(reinterpret_cast<const float4*>(ada_w + r * 512) + j * 32 + lane);
  const float bias = ada_b[r];
  for (int s = 0; s < ns; ++s) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 t = te[s][j * 32 + lane];
      acc += (w[j].x * t.x + w[j].y * t.y) + (w[j].z * t.z + w[j].w * t.w);
    }
    acc = warp_sum(acc);
    if (lane == 0) mod[(int64_t)(s0 + s) * R + r] = acc + bias;
  }
}

int dit_mod_table(const float* sigma, int S, const float* freqs, int half, const float* map0_w, const float* map0_b,
                  const float* map1_w, const float* map1_b, const float* ada_w, const float* ada_b, int depth,
                  int dim, float* t_emb_ws, float* mod, cudaStream_t stream) {
  RALD_REQUIRE(dim == 512, "dit_mod_table: dim=%d unsupported (512 only)", dim);
  RALD_REQUIRE(S > 0 && depth > 0, "dit_mod_table: bad sizes");
  const int smem = (2 * half + dim) * sizeof(float);
  temb_kernel<<<S, 512, smem, stream>>>(sigma, freqs, half, map0_w, map0_b, map1_w, map1_b, dim, t_emb_ws);
  RALD_LAUNCHED();
  const int64_t R = (int64_t)depth * 3 * 2 * dim;
  dim3 grid((unsigned)((R + 7) / 8), (unsigned)((S + ADA_SCHUNK - 1) / ADA_SCHUNK));
  adaln_kernel<<<grid, 256, 0, stream>>>(t_emb_ws, S, ada_w, ada_b, R, mod);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// boundary kernel: one warp per latent row (dim = 512, channels <= 32)
// ---------------------------------------------------------------------------------------------------
struct BoundaryParams {
  const float* h;          // [T, 512] residual stream after the last block (null when mode == 3)
  const float* ln_w;       // [512]
  const float* ln_b;       // [512]
  const float* w_out_t;    // [512][32]  proj_out transposed, zero-padded to 32 channels
  const float* w_in_t;     // [C][512]   proj_in transposed
  const float* x_in;       // [T, C] un-scaled network input of this evaluation
  const float* x_base;     // [T, C] x_hat of the current step (mode 2)
  float* d_buf;            // [T, C] d_cur (written in mode 1, read in mode 2)
  float* x_out;            // [T, C]
  float* h_next;           // [T, 512] or null
  const float* sigma;      // sigma of this evaluation (per frame with stride, or shared with stride 0)
  const float* sigma_other;
  int64_t sigma_stride, sigma_other_stride;
  int mode;                // 0 = D only, 1 = Euler, 2 = Heun, 3 = init (x_out = x_in * sigma), 4 = project only
  int rows_per_frame, C;
  int64_t T;
  float sigma_data;
};

// R rows per warp iteration. R = 1 (16 warps): 16 rows in flight per SM, the latency-bound small-batch form. R = 4
// (8 warps, 32 rows in flight): every weight value fetched from shared memory serves four rows — with one row per warp
// the two projections read 128 KB of shared memory per row and the kernel ran at the shared-memory bandwidth
// (194 us per 32 768 rows). The per-row arithmetic and its order are the same for both forms (bit-identical results).
template <int R, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
boundary_kernel(const BoundaryParams p) {
  extern __shared__ float sm[];
  float* s_wout = sm;                    // [512][32]
  float* s_win = s_wout + 512 * 32;      // [C][512] (allocated for 32 channels)
  float* s_row = s_win + 32 * 512;       // [WARPS][R][512]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool need_net = p.mode < 3;
  const bool need_next = p.h_next != nullptr;
  if (need_net) {
    for (int i = threadIdx.x; i < 512 * 32 / 4; i += blockDim.x)
      reinterpret_cast<float4*>(s_wout)[i] = __ldg(reinterpret_cast<const float4*>(p.w_out_t) + i);
  }
  if (need_next) {
    for (int i = threadIdx.x; i < p.C * 512 / 4; i += blockDim.x)
      reinterpret_cast<float4*>(s_win)[i] = __ldg(reinterpret_cast<const float4*>(p.w_in_t) + i);
  }
  pdl_wait();  // the weights staged above are constants; h / x / d come from the preceding kernels
  pdl_launch_dependents();
  __syncthreads();
  float* my_rows = s_row + warp * (R * 512);
  const float sd = p.sigma_data;

  for (int64_t row0 = ((int64_t)blockIdx.x * WARPS + warp) * R; row0 < p.T; row0 += (int64_t)gridDim.x * WARPS * R) {
    // the R rows of an iteration belong to one frame (R divides rows_per_frame, checked on the host)
    const int64_t f = row0 / p.rows_per_frame;
    const float sig = p.sigma[f * p.sigma_stride];
    const float sig_o = p.sigma_other ? p.sigma_other[f * p.sigma_other_stride] : 0.f;
    float F[R];
#pragma unroll
    for (int r = 0; r < R; ++r) F[r] = 0.f;
    if (need_net) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r;
        float4 v[4];
        if (row < p.T) {
          const float4* hr = reinterpret_cast<const float4*>(p.h + row * 512);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = hr[j * 32 + lane];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        const float mean = warp_sum(s) * (1.0f / 512);
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
          ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / 512) + 1e-5f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_w) + j * 32 + lane);
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.ln_b) + j * 32 + lane);
          reinterpret_cast<float4*>(my_rows + r * 512)[j * 32 + lane] =
              make_float4(v[j].x * rstd * g.x + b.x, v[j].y * rstd * g.y + b.y, v[j].z * rstd * g.z + b.z,
                          v[j].w * rstd * g.w + b.w);
        }
      }
      __syncwarp();
      float a[R][4];
#pragma unroll
      for (int r = 0; r < R; ++r) a[r][0] = a[r][1] = a[r][2] = a[r][3] = 0.f;
#pragma unroll 2
      for (int k = 0; k < 512; k += 4) {
        const float w0 = s_wout[(k + 0) * 32 + lane], w1 = s_wout[(k + 1) * 32 + lane];
        const float w2 = s_wout[(k + 2) * 32 + lane], w3 = s_wout[(k + 3) * 32 + lane];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 rv = *reinterpret_cast<const float4*>(my_rows + r * 512 + k);  // broadcast
          a[r][0] = fmaf(rv.x, w0, a[r][0]);
          a[r][1] = fmaf(rv.y, w1, a[r][1]);
          a[r][2] = fmaf(rv.z, w2, a[r][2]);
          a[r][3] = fmaf(rv.w, w3, a[r][3]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) F[r] = (a[r][0] + a[r][1]) + (a[r][2] + a[r][3]);
      __syncwarp();
    }
    const bool ch_ok = lane < p.C;
    float xs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = row0 + r;
      const bool ok = ch_ok && row < p.T;
      const float x = ok ? p.x_in[row * p.C + lane] : 0.f;
      float xo;
      float sig_next = sig;  // sigma at which the NEXT evaluation runs
      if (p.mode == 3) {
        xo = x * sig;   // x_0 = latents * t_0
      } else if (p.mode == 4) {
        xo = x;         // plain forward(): only the next projection h = proj_in(c_in x) is wanted
      } else {
        const float c_skip = sd * sd / (sig * sig + sd * sd);
        const float c_out = sig * sd / sqrtf(sig * sig + sd * sd);
        const float D = c_skip * x + c_out * F[r];
        if (p.mode == 0) {
          xo = D;
        } else if (p.mode == 1) {
          const float d = (x - D) / sig;
          xo = x + (sig_o - sig) * d;
          if (ok) p.d_buf[row * p.C + lane] = d;
          sig_next = sig_o;
        } else {
          const float dp = (x - D) / sig;
          const float dc = ok ? p.d_buf[row * p.C + lane] : 0.f;
          const float xb = ok ? p.x_base[row * p.C + lane] : 0.f;
          xo = xb + (sig - sig_o) * (0.5f * dc + 0.5f * dp);
        }
      }
      if (ok && p.x_out != nullptr) p.x_out[row * p.C + lane] = xo;
      const float c_in = 1.0f / sqrtf(sd * sd + sig_next * sig_next);
      xs[r] = ok ? c_in * xo : 0.f;
    }
    if (need_next) {
      float4 acc[R][4];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < p.C; ++c) {
        float xc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) xc[r] = __shfl_sync(0xffffffffu, xs[r], c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 w = reinterpret_cast<const float4*>(s_win + c * 512)[j * 32 + lane];
#pragma unroll
          for (int r = 0; r < R; ++r) {
            acc[r][j].x = fmaf(xc[r], w.x, acc[r][j].x);
            acc[r][j].y = fmaf(xc[r], w.y, acc[r][j].y);
            acc[r][j].z = fmaf(xc[r], w.z, acc[r][j].z);
            acc[r][j].w = fmaf(xc[r], w.w, acc[r][j].w);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (row0 + r < p.T) {
          float4* hn = reinterpret_cast<float4*>(p.h_next + (row0 + r) * 512);
#pragma unroll
          for (int j = 0; j < 4; ++j) hn[j * 32 + lane] = acc[r][j];
        }
      }
    }
  }
}

template <int R, int WARPS>
static int launch_boundary(const BoundaryParams& p, int64_t T, cudaStream_t stream) {
  const int smem = (512 * 32 + 32 * 512 + WARPS * R * 512) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(boundary_kernel<R, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int64_t blocks = (T + WARPS * R - 1) / (WARPS * R);
  const int64_t cap = device_sm_count();
  if (blocks > cap) blocks = cap;
  RALD_CHECK_CUDA(launch_pdl(boundary_kernel<R, WARPS>, dim3((unsigned)blocks), dim3(WARPS * 32), smem, stream, p));
  return 0;
}

int dit_boundary(const float* h, const float* ln_w, const float* ln_b, const float* w_out_t, const float* w_in_t,
                 const float* x_in, const float* x_base, float* d_buf, float* x_out, float* h_next,
                 const float* sigma, int64_t sigma_stride, const float* sigma_other, int64_t sigma_other_stride,
                 int mode, int rows_per_frame, int C, int64_t T, int dim, float sigma_data, cudaStream_t stream) {
  RALD_REQUIRE(dim == 512, "dit_boundary: dim=%d unsupported (512 only)", dim);
  RALD_REQUIRE(C >= 1 && C <= 32, "dit_boundary: channels=%d must be in [1, 32]", C);
  RALD_REQUIRE(mode >= 0 && mode <= 4, "dit_boundary: mode %d", mode);
  RALD_REQUIRE(mode >= 3 || h != nullptr, "dit_boundary: h missing");
  BoundaryParams p;
  p.h = h; p.ln_w = ln_w; p.ln_b = ln_b; p.w_out_t = w_out_t; p.w_in_t = w_in_t; p.x_in = x_in; p.x_base = x_base;
  p.d_buf = d_buf; p.x_out = x_out; p.h_next = h_next; p.sigma = sigma; p.sigma_other = sigma_other;
  p.sigma_stride = sigma_stride; p.sigma_other_stride = sigma_other_stride; p.mode = mode;
  p.rows_per_frame = rows_per_frame; p.C = C; p.T = T; p.sigma_data = sigma_data;
  ProfScope prof(FAM_BOUNDARY, stream, (double)T * (mode < 3 ? 2048.0 : 0.0) + (h_next ? (double)T * 2048.0 : 0.0) +
                                          (double)T * C * 16.0);
  // four rows per warp once the batch fills the machine with them (RALD_B200_BOUNDARY_R4=0: always one row per warp)
  static const bool r4_env = [] { const char* e = getenv("RALD_B200_BOUNDARY_R4"); return e == nullptr || e[0] != '0'; }();
  if (r4_env && rows_per_frame % 4 == 0 && T >= (int64_t)device_sm_count() * 8 * 4) RALD_TRY((launch_boundary<4, 8>(p, T, stream)));
  else RALD_TRY((launch_boundary<1, 16>(p, T, stream)));
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// radar tokens: tok[b, (r a e), :] = W x[b, r, a, e, :] + bias + r_emb[r] + a_emb[a] + e_emb[e]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
radar_tokens_kernel(const float* __restrict__ feat, int cz, int nr, int na, int ne, const float* __restrict__ w,
                    const float* __restrict__ b, const float* __restrict__ r_emb, const float* __restrict__ a_emb,
                    const float* __restrict__ e_emb, int dim, float* __restrict__ tok_f32,
                    __nv_bfloat16* __restrict__ tok_bf16) {
  const int64_t t = blockIdx.x;  // global token index over B * nr * na * ne
  const int e = (int)(t % ne);
  const int a = (int)((t / ne) % na);
  const int r = (int)((t / ((int64_t)ne * na)) % nr);
  const float* x = feat + t * cz;
  for (int j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < cz; ++c) acc = fmaf(w[(int64_t)j * cz + c], x[c], acc);
    acc += b[j];
    acc = ((acc + r_emb[(int64_t)r * dim + j]) + a_emb[(int64_t)a * dim + j]) + e_emb[(int64_t)e * dim + j];
    if (tok_f32) tok_f32[t * dim + j] = acc;
    if (tok_bf16) tok_bf16[t * dim + j] = __float2bfloat16_rn(acc);
  }
}

int radar_tokens(const float* feat, int B, int nr, int na, int ne, int cz, const float* w, const float* b,
                 const float* r_emb, const float* a_emb, const float* e_emb, int dim, float* tok_f32, void* tok_bf16,
                 cudaStream_t stream) {
  const int64_t ntok = (int64_t)B * nr * na * ne;
  RALD_REQUIRE(ntok > 0 && ntok < (1ll << 31), "radar_tokens: bad token count");
  radar_tokens_kernel<<<(unsigned)ntok, 128, 0, stream>>>(feat, cz, nr, na, ne, w, b, r_emb, a_emb, e_emb, dim,
                                                          tok_f32, reinterpret_cast<__nv_bfloat16*>(tok_bf16));
  RALD_LAUNCHED();
  return 0;
}

}  // namespace rald
