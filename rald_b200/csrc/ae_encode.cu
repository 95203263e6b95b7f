// VecSet autoencoder, encode side (KLAutoEncoder.encode, model/models_ae.py:351-405):
//   point_features_kernel  [sin(p f), cos(p f), p] -> bf16 [rows, 64] (51 features zero padded)       PointEmbed :128-137
//   fps_kernel             farthest point sampling, one CTA per cloud, points staged in shared memory    torch_cluster.fps, :359-370
//   softmax_rows_kernel    fp32 scores [rows, ld] -> bf16 probabilities, columns >= n written as 0       Attention :91-101
//   posterior_kernel       clamp(logvar), z = mean + exp(logvar/2) noise, KL per frame                   :141-163
//   rald_ae_encode_stats   the whole encode up to (mean, logvar) as a sequence of launches
//
// The two long-context attentions (512 queries against N = 10000 point embeddings; 8 x 64 heads for the "mix"
// query path, ONE head of width 512 for cross_attend_blocks[0]) are run per frame as
//   S = Q K^T (tcgen05 GEMM, fp32 scores) -> row softmax (bf16 P) -> O = P V (tcgen05 GEMM against V^T),
// where V^T comes straight out of the projection GEMM (A = W_v, W = context), so no transpose pass exists.
// A head width of 512 does not fit a TMEM-resident flash tile (O alone is 128 x 512 fp32 = all 512 columns).
#include <cuda_bf16.h>
#include <math.h>

#include "../../include/rald_b200.h"
#include "host.cuh"
#include "kernels.h"
#include "ptx.cuh"

namespace rald {

// ---------------------------------------------------------------------------------------------------
// Fourier point features. idx (optional): gather rows pts[b][idx[b][i]] (the FPS-selected points).
// ---------------------------------------------------------------------------------------------------
struct FeatParams {
  const float* pts;        // [B, N, 3]
  const int64_t* idx;      // [B, M] indices into each cloud, or null
  __nv_bfloat16* feat;     // [rows, 64]
  float* gathered;         // [B, M, 3] optional copy of the gathered points
  int64_t rows;            // B*N, or B*M with idx
  int64_t N, M;
  float freq[3][8];
};

__global__ void __launch_bounds__(256)
point_features_kernel(const FeatParams p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.rows) return;
  const float* src;
  if (p.idx != nullptr) {
    const int64_t b = r / p.M;
    src = p.pts + (b * p.N + p.idx[r]) * 3;
  } else {
    src = p.pts + r * 3;
  }
  const float pt[3] = {src[0], src[1], src[2]};
  if (p.gathered != nullptr) {
    p.gathered[r * 3 + 0] = pt[0]; p.gathered[r * 3 + 1] = pt[1]; p.gathered[r * 3 + 2] = pt[2];
  }
  float f[64];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float sn, cs;
      sincosf(pt[a] * p.freq[a][k], &sn, &cs);
      f[a * 8 + k] = sn;
      f[24 + a * 8 + k] = cs;
    }
  }
  f[48] = pt[0]; f[49] = pt[1]; f[50] = pt[2];
#pragma unroll
  for (int j = 51; j < 64; ++j) f[j] = 0.f;
  uint4* dst = reinterpret_cast<uint4*>(p.feat + r * 64);
#pragma unroll
  for (int c = 0; c < 8; ++c)
    dst[c] = make_uint4(pack_bf16x2(f[8 * c + 0], f[8 * c + 1]), pack_bf16x2(f[8 * c + 2], f[8 * c + 3]),
                        pack_bf16x2(f[8 * c + 4], f[8 * c + 5]), pack_bf16x2(f[8 * c + 6], f[8 * c + 7]));
}

int point_features(const float* pts, const int64_t* idx, int B, int64_t N, int64_t M, const float* freq24,
                   void* feat_bf16, float* gathered, cudaStream_t stream) {
  FeatParams p;
  p.pts = pts; p.idx = idx; p.feat = reinterpret_cast<__nv_bfloat16*>(feat_bf16); p.gathered = gathered;
  p.N = N; p.M = M; p.rows = (int64_t)B * (idx ? M : N);
  for (int a = 0; a < 3; ++a)
    for (int k = 0; k < 8; ++k) p.freq[a][k] = freq24[a * 8 + k];
  point_features_kernel<<<(unsigned)((p.rows + 255) / 256), 256, 0, stream>>>(p);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Farthest point sampling. One CTA of 1024 threads per cloud; the cloud lives in shared memory as SoA
// (x[], y[], z[], each padded to a multiple of 4 points), thread t owns the point quads [4 t + 4096 j, +4) and keeps
// their running minimum distances in registers. Per pick (511 dependent picks for M = 512):
//   sweep      3 x LDS.128 per quad; differences and squares as packed fp32x2 operations (FADD2 / FMUL2: two points
//              per instruction, each lane individually rounded), the two sums as scalar FADDs — ptxas would contract a
//              packed multiply + packed add into FFMA2 even with .rn, which the definition below forbids;
//   argmax     thread-local (value, index) -> per warp TWO redux.sync (max of the value's bit pattern, then min index
//              among the lanes holding it) -> 32 (value, index) pairs through double-buffered shared memory and ONE
//              barrier -> every warp reduces the 32 pairs itself the same way (no second barrier, no broadcast).
// Round 1's form (scalar arithmetic, 10 shuffles per level, two barriers) took 1.6 us per pick for N = 10000.
// Deterministic definition (oracle/fps_oracle.c, oracle/rald_oracle.py:fps_indices): start at index 0; squared
// distance (dx*dx + dy*dy) + dz*dz with individually rounded fp32 operations; ties -> lowest index.
// ---------------------------------------------------------------------------------------------------
constexpr int FPS_THREADS = 1024;
constexpr int FPS_QUADS = 4;         // quads per thread: clouds up to 4 * 4 * 1024 = 16384 points
constexpr int FPS_MAX_POINTS = FPS_THREADS * FPS_QUADS * 4;

__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// (value, index) argmax over the 32 lanes of a warp, ties -> lowest index. The values are >= 0 or -inf, so the signed
// integer order of their bit patterns is the float order.
__device__ __forceinline__ void warp_argmax(int& key, int& idx) {
  const int kmax = __reduce_max_sync(0xffffffffu, key);
  idx = (int)__reduce_min_sync(0xffffffffu, key == kmax ? (unsigned)idx : 0xffffffffu);
  key = kmax;
}

__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_kernel(const float* __restrict__ pts, int N, int M, int64_t* __restrict__ out) {
  extern __shared__ __align__(16) float sm[];
  const int Np = (N + 3) & ~3;
  float* sx = sm;
  float* sy = sx + Np;
  float* sz = sy + Np;
  __shared__ int s_key[2][32];
  __shared__ int s_idx[2][32];
  const int b = blockIdx.x;
  const float* p = pts + (int64_t)b * N * 3;
  for (int i = threadIdx.x; i < Np; i += FPS_THREADS) {
    const bool ok = i < N;
    sx[i] = ok ? p[3 * i + 0] : 0.f; sy[i] = ok ? p[3 * i + 1] : 0.f; sz[i] = ok ? p[3 * i + 2] : 0.f;
  }
  // running minimum distance: +inf for real points, -inf for the padding (min keeps it there, so it never wins)
  float dist[FPS_QUADS][4];
#pragma unroll
  for (int j = 0; j < FPS_QUADS; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) dist[j][e] = (4 * (int)threadIdx.x + j * 4 * FPS_THREADS + e) < N ? INFINITY : -INFINITY;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cur = 0;
  for (int m = 0; m < M; ++m) {
    if (threadIdx.x == 0) out[(int64_t)b * M + m] = cur;
    if (m == M - 1) break;
    const float cx = sx[cur], cy = sy[cur], cz = sz[cur];
    const unsigned long long cx2 = f2_pack(cx, cx), cy2 = f2_pack(cy, cy), cz2 = f2_pack(cz, cz);
    float bd = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < FPS_QUADS; ++j) {
      const int i0 = 4 * (int)threadIdx.x + j * 4 * FPS_THREADS;   // ascending index per thread: '>' keeps the lowest
      if (i0 < Np) {                                               // index on ties
        const float4 vx = *reinterpret_cast<const float4*>(sx + i0);
        const float4 vy = *reinterpret_cast<const float4*>(sy + i0);
        const float4 vz = *reinterpret_cast<const float4*>(sz + i0);
        float d2[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const unsigned long long dx = f2_sub(h ? f2_pack(vx.z, vx.w) : f2_pack(vx.x, vx.y), cx2);
          const unsigned long long dy = f2_sub(h ? f2_pack(vy.z, vy.w) : f2_pack(vy.x, vy.y), cy2);
          const unsigned long long dz = f2_sub(h ? f2_pack(vz.z, vz.w) : f2_pack(vz.x, vz.y), cz2);
          float xa, xb, ya, yb, za, zb;
          f2_unpack(f2_mul(dx, dx), xa, xb);
          f2_unpack(f2_mul(dy, dy), ya, yb);
          f2_unpack(f2_mul(dz, dz), za, zb);
          d2[2 * h] = __fadd_rn(__fadd_rn(xa, ya), za);
          d2[2 * h + 1] = __fadd_rn(__fadd_rn(xb, yb), zb);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = fminf(dist[j][e], d2[e]);
          dist[j][e] = d;
          if (d > bd) { bd = d; bi = i0 + e; }
        }
      }
    }
    int key = __float_as_int(bd);
    warp_argmax(key, bi);
    const int buf = m & 1;
    if (lane == 0) { s_key[buf][warp] = key; s_idx[buf][warp] = bi; }
    __syncthreads();
    key = s_key[buf][lane];
    bi = s_idx[buf][lane];
    warp_argmax(key, bi);
    cur = bi;
  }
}

int fps(const float* pts, int B, int N, int M, int64_t* out_idx, cudaStream_t stream) {
  RALD_REQUIRE(B > 0 && N > 0 && M > 0, "fps: bad sizes B=%d N=%d M=%d", B, N, M);
  RALD_REQUIRE(N <= FPS_MAX_POINTS, "fps: N=%d exceeds the %d points one CTA holds", N, FPS_MAX_POINTS);
  RALD_REQUIRE(M <= N, "fps: cannot sample %d of %d points", M, N);
  const int smem = 3 * ((N + 3) & ~3) * sizeof(float);
  RALD_REQUIRE(smem <= 200 * 1024, "fps: cloud of %d points does not fit in shared memory", N);
  // the kernel also has static shared memory, so even a request of exactly 48 KB needs the opt-in
  static int configured = 32 * 1024;
  if (smem > configured) {
    RALD_CHECK_CUDA(cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
  ProfScope prof(FAM_FPS, stream, (double)B * N * (double)M * 12.0);  // shared-memory bytes swept
  fps_kernel<<<B, FPS_THREADS, smem, stream>>>(pts, N, M, out_idx);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Row softmax: P[r][j] = exp(scale (S[r][j] - max_j S[r][j])) / sum for j < n, 0 for n <= j < ld_out.
// One CTA of 256 threads per row.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, int64_t lds, int n, float scale, __nv_bfloat16* __restrict__ P,
                    int64_t ldp, int n_pad) {
  __shared__ float s_red[8];
  __shared__ float s_bc;
  const float* s = S + (int64_t)blockIdx.x * lds;
  __nv_bfloat16* p = P + (int64_t)blockIdx.x * ldp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, s[j]);
  mx = warp_max(mx);
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = s_red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i]);
    s_bc = m;
  }
  __syncthreads();
  mx = s_bc;
  float sum = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) sum += expf(scale * (s[j] - mx));
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) s_red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s_red[i];
    s_bc = 1.0f / t;
  }
  __syncthreads();
  const float inv = s_bc;
  for (int j = threadIdx.x; j < n_pad; j += 256)
    p[j] = __float2bfloat16_rn(j < n ? expf(scale * (s[j] - mx)) * inv : 0.f);
}

int softmax_rows(const float* S, int64_t lds, int rows, int n, float scale, void* P_bf16, int64_t ldp, int n_pad,
                 cudaStream_t stream) {
  ProfScope prof(FAM_OTHER, stream, (double)rows * n * 10.0);
  softmax_rows_kernel<<<rows, 256, 0, stream>>>(S, lds, n, scale, reinterpret_cast<__nv_bfloat16*>(P_bf16), ldp,
                                                n_pad);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Posterior: ml [T, ld] = (mean | logvar), noise [T, C] -> logvar clamped to [-30, 20] (written back),
// z = mean + exp(logvar / 2) noise, kl[b] = 0.5 mean(mean^2 + exp(logvar) - 1 - logvar). One CTA per frame.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
posterior_kernel(const float* __restrict__ ml, int64_t ld, const float* __restrict__ noise, int rows_per_frame, int C,
                 float* __restrict__ mean, float* __restrict__ logvar, float* __restrict__ z, float* __restrict__ kl) {
  __shared__ double s_red[8];
  const int b = blockIdx.x;
  const int64_t n = (int64_t)rows_per_frame * C;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) {
    const int64_t row = (int64_t)b * rows_per_frame + i / C;
    const int c = (int)(i % C);
    const float m = ml[row * ld + c];
    float lv = ml[row * ld + C + c];
    lv = fminf(fmaxf(lv, -30.0f), 20.0f);
    const int64_t o = row * C + c;
    mean[o] = m;
    logvar[o] = lv;
    if (z != nullptr) z[o] = m + expf(0.5f * lv) * noise[o];
    acc += (double)(m * m + expf(lv) - 1.0f - lv);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s_red[i];
    kl[b] = (float)(0.5 * t / (double)n);
  }
}

int ae_posterior(const float* ml, int64_t ld, const float* noise, int B, int rows_per_frame, int C, float* mean,
                 float* logvar, float* z, float* kl, cudaStream_t stream) {
  RALD_REQUIRE(z == nullptr || noise != nullptr, "ae_posterior: sampling needs the noise tensor");
  posterior_kernel<<<B, 256, 0, stream>>>(ml, ld, noise, rows_per_frame, C, mean, logvar, z, kl);
  RALD_LAUNCHED();
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// Long-context attention of ONE frame through GEMMs: out[Sq, heads*dh] = softmax(q k^T / sqrt(dh)) v per head.
//   q   bf16 [Sq, heads*dh] (ldq);  k bf16 [n, heads*dh] (ldk);  vt bf16 [heads*dh, n_pad] (V transposed)
//   scores fp32 scratch [Sq, n_pad]; prob bf16 scratch [Sq, n_pad]
// ---------------------------------------------------------------------------------------------------
static int long_attention(const __nv_bfloat16* q, int64_t ldq, const __nv_bfloat16* k, int64_t ldk,
                          const __nv_bfloat16* vt, int n, int n_pad, int Sq, int heads, int dh, float* scores,
                          __nv_bfloat16* prob, __nv_bfloat16* out, int64_t ldo, cudaStream_t st) {
  const float scale = 1.0f / sqrtf((float)dh);
  for (int h = 0; h < heads; ++h) {
    // S = q_h k_h^T : "W" = k rows (n real rows; TMA zero-fills rows n..n_pad)
    RALD_TRY(gemm_bf16(q + h * dh, ldq, k + h * dh, ldk, scores, n_pad, nullptr, nullptr, 0, Sq, n_pad, dh, 1, 0, st));
    RALD_TRY(softmax_rows(scores, n_pad, Sq, n, scale, prob, n_pad, n_pad, st));
    // O_h = P V_h : "W" = rows [h*dh, (h+1)*dh) of V^T, reduction over n_pad keys (padded P columns are 0)
    RALD_TRY(gemm_bf16(prob, n_pad, vt + (int64_t)h * dh * n_pad, n_pad, out + h * dh, ldo, nullptr, nullptr, 0, Sq, dh,
                       n_pad, 0, 0, st));
  }
  return 0;
}

}  // namespace rald

using namespace rald;

extern "C" int rald_point_features(const float* pts, const int64_t* idx, int B, int64_t N, int64_t M,
                                   const float* freq24_host, void* feat_bf16, float* gathered, void* stream) {
  return point_features(pts, idx, B, N, M, freq24_host, feat_bf16, gathered, static_cast<cudaStream_t>(stream));
}

extern "C" int rald_fps(const float* pts, int B, int N, int M, int64_t* out_idx, void* stream) {
  return fps(pts, B, N, M, out_idx, static_cast<cudaStream_t>(stream));
}

extern "C" int rald_ae_posterior(const float* ml, int64_t ld, const float* noise, int B, int rows_per_frame, int C,
                                 float* mean, float* logvar, float* z, float* kl, void* stream) {
  return ae_posterior(ml, ld, noise, B, rows_per_frame, C, mean, logvar, z, kl, static_cast<cudaStream_t>(stream));
}

extern "C" int rald_ae_encode_stats(const rald_ae_enc_weights* w, const rald_ae_enc_workspace* ws, const float* pc,
                                    int B, int N, int64_t* fps_idx, float* ml_out, void* stream) {
  RALD_REQUIRE(w != nullptr && ws != nullptr, "ae_encode: null weights/workspace");
  RALD_REQUIRE(w->dim == 512 && w->n_latents % 128 == 0, "ae_encode: dim=%d latents=%d unsupported", w->dim,
               w->n_latents);
  RALD_REQUIRE(w->query_type >= 0 && w->query_type <= 2, "ae_encode: query_type %d", w->query_type);
  RALD_REQUIRE(N >= 32 && N <= ws->max_points, "ae_encode: N=%d outside the workspace (max %d)", N, ws->max_points);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dim = w->dim, M = w->n_latents;
  const int n_pad = (N + 31) / 32 * 32;
  __nv_bfloat16* feat = reinterpret_cast<__nv_bfloat16*>(ws->feat);
  __nv_bfloat16* pe16 = reinterpret_cast<__nv_bfloat16*>(ws->pe16);
  __nv_bfloat16* kbuf = reinterpret_cast<__nv_bfloat16*>(ws->kbuf);
  __nv_bfloat16* vt = reinterpret_cast<__nv_bfloat16*>(ws->vt);
  __nv_bfloat16* prob = reinterpret_cast<__nv_bfloat16*>(ws->prob);
  __nv_bfloat16* xq = reinterpret_cast<__nv_bfloat16*>(ws->xq);
  __nv_bfloat16* att = reinterpret_cast<__nv_bfloat16*>(ws->att);
  __nv_bfloat16* xn = reinterpret_cast<__nv_bfloat16*>(ws->xn);
  __nv_bfloat16* ff = reinterpret_cast<__nv_bfloat16*>(ws->ff);
  const int n_stats = w->stats_rows;  // 2*latent_dim rounded up to 32
  for (int b = 0; b < B; ++b) {
    const float* pts = pc + (int64_t)b * N * 3;
    float* x = ws->x;  // [M, dim] fp32 query / residual stream of this frame
    // ---- point embeddings of the whole cloud: pe = Linear(51 -> dim)(features) ----
    RALD_TRY(point_features(pts, nullptr, 1, N, 0, w->freq24, feat, nullptr, st));
    RALD_TRY(gemm_bf16(feat, 64, w->wpe, 64, ws->pe32, dim, w->pe_bias, nullptr, 0, N, dim, 64, 1, 0, st));
    // ---- initial queries ----
    if (w->query_type == 0) {
      // FPS -> point_embed(sampled points) (:357-377)
      int64_t* idx = fps_idx + (int64_t)b * M;
      RALD_TRY(fps(pts, 1, N, M, idx, st));
      RALD_TRY(point_features(pts, idx, 1, N, M, w->freq24, xq, nullptr, st));
      RALD_TRY(gemm_bf16(xq, 64, w->wpe, 64, x, dim, w->pe_bias, nullptr, 0, M, dim, 64, 1, 0, st));
    } else if (w->query_type == 1) {
      RALD_CHECK_CUDA(cudaMemcpyAsync(x, w->latents, sizeof(float) * M * dim, cudaMemcpyDeviceToDevice, st));
    } else {
      // mix (:378-387): dynamic = mix_attn(LN(d_latents), context = RAW point embeddings); x = query_proj(static + dynamic)
      RALD_TRY(gemm_bf16(feat, 64, w->wpe, 64, pe16, dim, w->pe_bias, nullptr, 0, N, dim, 64, 0, 0, st));
      const __nv_bfloat16* wkv = reinterpret_cast<const __nv_bfloat16*>(w->mix_wkv);
      RALD_TRY(gemm_bf16(pe16, dim, wkv, dim, kbuf, dim, nullptr, nullptr, 0, N, dim, dim, 0, 0, st));
      RALD_TRY(gemm_bf16(wkv + (int64_t)dim * dim, dim, pe16, dim, vt, n_pad, nullptr, nullptr, 0, dim, n_pad, dim, 0, 0,
                         st));
      RALD_TRY(long_attention(reinterpret_cast<const __nv_bfloat16*>(w->mix_q), dim, kbuf, dim, vt, N, n_pad, M,
                              w->heads, dim / w->heads, ws->scores, prob, att, dim, st));
      // (static + dynamic) in the epilogue: out bf16 = att W_o^T + b_o + s_latents
      RALD_TRY(gemm_bf16_ex(att, dim, w->mix_wo, dim, xq, dim, w->mix_bo, w->s_latents, dim, M, M, dim, dim, 0, 0, st));
      RALD_TRY(gemm_bf16(xq, dim, w->wproj, dim, x, dim, w->bproj, nullptr, 0, M, dim, dim, 1, 0, st));
    }
    // ---- x += cross_attn(LN(x), LN_ctx(pe))  (1 head of width dim, :392-395) ----
    RALD_TRY(ln_rows(ws->pe32, dim, w->ca_lnc_w, w->ca_lnc_b, 0, 0, 0, pe16, dim, 0, N, dim, 1e-5f, st));
    RALD_TRY(ln_rows(x, dim, w->ca_ln_w, w->ca_ln_b, 0, 0, 0, xn, dim, 0, M, dim, 1e-5f, st));
    RALD_TRY(gemm_bf16(xn, dim, w->ca_wq, dim, xq, dim, nullptr, nullptr, 0, M, dim, dim, 0, 0, st));
    const __nv_bfloat16* wkv = reinterpret_cast<const __nv_bfloat16*>(w->ca_wkv);
    RALD_TRY(gemm_bf16(pe16, dim, wkv, dim, kbuf, dim, nullptr, nullptr, 0, N, dim, dim, 0, 0, st));
    RALD_TRY(gemm_bf16(wkv + (int64_t)dim * dim, dim, pe16, dim, vt, n_pad, nullptr, nullptr, 0, dim, n_pad, dim, 0, 0,
                       st));
    RALD_TRY(long_attention(xq, dim, kbuf, dim, vt, N, n_pad, M, 1, dim, ws->scores, prob, att, dim, st));
    RALD_TRY(gemm_bf16(att, dim, w->ca_wo, dim, x, dim, w->ca_bo, x, dim, M, dim, dim, 1, 0, st));
    // ---- x += FF(LN(x)) (:396) ----
    RALD_TRY(ln_rows(x, dim, w->ff_ln_w, w->ff_ln_b, 0, 0, 0, xn, dim, 0, M, dim, 1e-5f, st));
    RALD_TRY(gemm_bf16(xn, dim, w->ff_w1, dim, ff, 4 * dim, w->ff_b1, nullptr, 0, M, 8 * dim, dim, 2, 0, st));
    RALD_TRY(gemm_bf16(ff, 4 * dim, w->ff_w2, 4 * dim, x, dim, w->ff_b2, x, dim, M, dim, 4 * dim, 1, 0, st));
    if (w->w_stats == nullptr) {
      // deterministic AutoEncoder.encode (:226-257): the latents are the fp32 residual stream itself, [M][dim]
      RALD_CHECK_CUDA(cudaMemcpyAsync(ml_out + (int64_t)b * M * dim, x, sizeof(float) * M * dim, cudaMemcpyDeviceToDevice, st));
      continue;
    }
    // ---- (mean | logvar) = x [W_mean; W_logvar]^T + b (:398-399), x taken in fp32 -> bf16 without a norm ----
    RALD_TRY(gn_apply(x, nullptr, nullptr, nullptr, xn, 1, (int64_t)M * dim / 256, 256, 0, 0.f, 2, st));
    RALD_TRY(gemm_bf16(xn, dim, w->w_stats, dim, ml_out + (int64_t)b * M * n_stats, n_stats, w->b_stats, nullptr, 0, M,
                       n_stats, dim, 1, 0, st));
  }
  return 0;
}
