/*
 * rald_b200 — C ABI of the B200 (sm_100a) kernels behind the RaLD generation hot path.
 *
 * The reference (RoyAPTX4869/RaLD) has no FFI layer: its hot path is Python nn.Modules calling torch ops
 * (SURVEY.md §8b). This header is the boundary a maintainer binds instead of those torch ops; the Python
 * modules in rald_b200/models_*.py bind it with ctypes. Each entry point cites the
 * reference code it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are dense row-major;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it;
 *   - the library never allocates or frees device memory and keeps no ownership of any buffer;
 *   - return value 0 = success, negative = error (rald_last_error() gives a thread-local message);
 *   - bf16 = 16-bit bfloat16, f32 = IEEE binary32, i64 = int64_t.
 */
#ifndef RALD_B200_H
#define RALD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RALD_ABI_VERSION 5

int rald_abi_version(void);
const char* rald_last_error(void);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t rald_launch_count(void);
/* Accounts for kernels launched by replaying a CUDA graph captured from this library's calls (the host runtime adds
 * the number of launches it counted during capture at every replay). */
void rald_launch_count_add(uint64_t n);
/* TMA descriptor cache of the library (keyed on pointer + geometry, mutex-guarded): encodes avoided /
 * performed since the process started. */
int rald_tmap_cache_stats(uint64_t* hits, uint64_t* misses);

/* Per-launch timing for bench.py's roofline leg: rald_prof_enable(mask) starts a session that brackets every launch
 * of the kernel families in `mask` (bit = family id below) with CUDA events on the launching stream (0 = off);
 * rald_prof_collect(family, ...) synchronises and returns the summed duration (ms), summed algorithmic work
 * (flops for tensor-core kernels, bytes for HBM-bound ones) and the launch count of one family. */
#define RALD_FAM_GEMM 0
#define RALD_FAM_ATTN 1
#define RALD_FAM_LN 2
#define RALD_FAM_BOUNDARY 3
#define RALD_FAM_CONV3D 4
#define RALD_FAM_GN 5
#define RALD_FAM_AE_QUERY 6
#define RALD_FAM_OTHER 7
#define RALD_FAM_FPS 8
#define RALD_FAM_XATTN 9
int rald_prof_enable(unsigned family_mask);
int rald_prof_collect(int family, double* total_ms, double* total_work, int64_t* launches);
/* Per-launch records of one family (HOST arrays ms[cap], work[cap]); returns how many were written, -1 on error. */
int64_t rald_prof_dump(int family, float* ms, double* work, int64_t cap);

/* out = epilogue(A[M,K] @ W[N,K]^T), A and W bf16, fp32 accumulation on tcgen05 tensor cores.
 *   out_mode 0: out bf16 [M,N] (+bias) (+resid f32, added before rounding)
 *   out_mode 1: out f32  [M,N] = acc (+bias) (+resid f32 [M,ldr]); out may alias resid
 *   out_mode 2: GEGLU, out bf16 [M,N/2] = value * gelu(gate); W rows / bias packed in groups of 32 = 16 value + 16
 *               gate rows. gelu is the erf GELU of F.gelu evaluated in logistic form x * sigmoid(x * P(x^2)) with
 *               |error| <= 2.5e-5 (build with -DRALD_GELU_LOGISTIC=0 for the A&S 7.1.26 erf polynomial, 1.5e-7)
 * Replaces nn.Linear (+ residual add / GEGLU) at model/models_radar_generation.py:58-64, 76, 91-95, 113,
 * 166-168 and model/models_ae.py:60-62, 87-89, 105. bn_hint: 0 = auto, else 32/64/128/256 (tile N). */
int rald_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                   const float* bias, const float* resid, int64_t ldr, int M, int N, int K, int out_mode,
                   int bn_hint, void* stream);

/* out (f32 [M,N], pitch ldo) += A[M,K] @ W[N,K]^T for GEMMs with few output tiles and a long K (the weight gradients
 * dW = dY^T X of the training step, K = rows of the batch): the K extent of every tile is split over up to 16 CTAs that
 * add into `out` through the TMA reduce-add epilogue. `out` must hold the value to accumulate onto (zeros for a fresh
 * gradient). The order of the fp32 additions across splits is not fixed: results reproduce to ~1e-7 relative. */
int rald_gemm_bf16_accum(const void* A, int64_t lda, const void* W, int64_t ldw, float* out, int64_t ldo, int M, int N,
                         int K, void* stream);

/* rald_gemm_bf16_accum with the W operand read w_col_shift columns to the right of A: out[m][n] += sum_k A[m][k] *
 * W[n][k + w_col_shift] (w_col_shift % 8 == 0), columns of W outside [0, K) counting as zero. One (kd, kh) tap pair of a 3x3x3 convolution's weight gradient
 * when A = dY^T and W = X^T are laid out on the flattened zero-padded voxel grid (rald_enc_pad_transpose), where a tap is
 * a constant index offset (autograd through nn.Conv3d, model/models_radar_encoder.py:63-72, 37-41). */
int rald_gemm_bf16_accum_shift(const void* A, int64_t lda, const void* W, int64_t ldw, int w_col_shift, float* out,
                               int64_t ldo, int M, int N, int K, void* stream);

/* n_taps (<= 16) shifted products of rald_gemm_bf16_accum_shift in ONE launch: out[m][t * w_rows + n] += sum_k A[m][k] *
 * W[n][k + tap_shifts_host[t]] with W of w_rows rows (a multiple of 32) — all (kd, kh) taps of a convolution's weight
 * gradient; the CTAs that work on different taps of the same K range run concurrently and share A and W in L2, so HBM
 * sees the operands about once instead of n_taps times; with w_rows of 64 or 128 several taps share one 256-wide tile
 * (one A tile, one MMA). tap_shifts_host: HOST array of n_taps multiples of 8. panel_len > 0 (a multiple of 64 that
 * divides K; lda / ldw unused): K-PANEL-MAJOR operands, A = [K/panel_len][M][panel_len] and W = [K/panel_len][w_rows]
 * [panel_len + 2*w_halo] where every W panel repeats w_halo columns of its neighbours on both sides (|shift| <= w_halo).
 * With row-major operands of several MB of row pitch every row of a TMA box lies in another 2 MB page and the same
 * product runs 3x slower (measured: 346 vs 1042 TFLOP/s); rald_enc_pad_transpose writes this layout. */
int rald_gemm_bf16_accum_taps(const void* A, int64_t lda, const void* W, int64_t ldw, int w_rows, int n_taps,
                              const int* tap_shifts_host, int panel_len, int w_halo, float* out, int64_t ldo, int M, int K,
                              void* stream);

/* rald_gemm_bf16 with out_mode 0 whose output columns with (col % f16_period) >= f16_start are written as IEEE fp16
 * instead of bf16 (start / period multiples of 64): the V projections consumed by rald_attn_d64. */
int rald_gemm_bf16_f16cols(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                           const float* bias, int M, int N, int K, int f16_start, int f16_period, void* stream);

/* Split-weight GEMM: W_hilo is bf16 [N][2*K] = [W_hi | W_lo] with W_hi = bf16(W), W_lo = bf16(W - W_hi) (16 mantissa
 * bits of the fp32 nn.Linear weight); out = epilogue(A W_hi^T + A W_lo^T) in ONE fp32 accumulator, the A tiles being
 * read twice. out_mode as rald_gemm_bf16; f16_period > 0 selects fp16 columns as rald_gemm_bf16_f16cols (0: none);
 * gelu_exact != 0: the GEGLU epilogue evaluates the erf GELU (A&S 7.1.26, |error| <= 1.5e-7). Used for the 24-layer
 * latent stack of KLAutoEncoder.decode (model/models_ae.py:410-414), whose bf16 weight rounding is otherwise a
 * common-mode error of the whole occupancy field. */
int rald_gemm_bf16_wsplit(const void* A, int64_t lda, const void* W_hilo, int64_t ldw, void* out, int64_t ldo,
                          const float* bias, const float* resid, int64_t ldr, int M, int N, int K, int out_mode,
                          int f16_start, int f16_period, int gelu_exact, void* stream);

/* Debug hook: when dev_buf != NULL every GEMM CTA stores %globaltimer stamps of its first tile at dev_buf[cta*8 + i]
 * (0 entry, 1 setup done, 2 first operands landed, 3 last MMA issued, 4 accumulator ready, 5 epilogue done, 6 exit). */
int rald_gemm_debug_buffer(unsigned long long* dev_buf);

/* Debug hook like rald_gemm_debug_buffer for the attention kernel: CTA 0 stores, for its first 16 query tiles and both
 * softmax groups, stamps at dev_buf[(tile*2+group)*8 + i]: 0 tile start, 1 S ready, 2 P written, 3 stats merged,
 * 4 O ready, 5 stored. */
int rald_attn_debug_buffer(unsigned long long* dev_buf);
/* Same for the four-stream form of the Skv = 512 case (attn_streams.cu): issuer and stream stamps of CTA 0. */
int rald_attn_streams_debug_buffer(unsigned long long* dev_buf);

/* O = softmax(Q K^T * scale) V per (frame, head), head_dim 64, Skv <= 512 (multiple of 64), scores kept in TMEM.
 * Q: [frames*Sq, >= heads*64] bf16 (ldq), K: [frames*Skv, ...] bf16, V: [frames*Skv, ...] **fp16** (the probabilities
 * are produced as fp16 by a packed exp2, and tcgen05 kind::f16 needs both operands of P V in one format),
 * O: [frames*Sq, ...] bf16; head h uses
 * columns [h*64, h*64+64). Replaces the two einsums + softmax of CrossAttention.forward
 * (model/models_radar_generation.py:66-75) and Attention.forward (model/models_ae.py:91-104). */
int rald_attn_d64(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                  int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, void* stream);

/* rald_attn_d64 for contexts longer than TMEM holds: Skv = chunks * 512 keys per frame (1024 .. 4096), processed as
 * key chunks with per-chunk softmax statistics and merged exactly (O = sum_c w_c O_c / sum_c w_c, w_c = l_c 2^(m_c - m)).
 * Serves `use_radar_enc: false` (2048 raw radar-cube tokens as the cross-attention context,
 * model/models_radar_generation.py:357-361, 378-405). Scratch: o_chunks bf16 [chunks][frames*Sq][heads*64],
 * stats f32 [chunks][frames*Sq][heads][2]. */
int rald_attn_d64_long(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                       int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, void* o_chunks, float* stats,
                       void* stream);

/* out = LN(x) * g + b over rows of 512 fp32 values (eps inside the rsqrt). gamma_plus_one=1 gives the adaLN
 * modulation LN(x)*(1+scale)+shift of AdaLayerNorm.forward (model/models_radar_generation.py:127-131) with
 * gamma/beta = scale/shift of frame f at gamma + f*mod_frame_stride (stride 0 = shared); gamma_plus_one=0 is
 * nn.LayerNorm with affine weights (model/models_ae.py:38-47). out_f32=0 writes bf16. gamma / beta are fetched BEFORE the
 * kernel's programmatic-dependency wait (they are constants of a sampling / training step): they must not be written by the
 * launch that immediately precedes this one on the stream (x may be). */
int rald_ln_rows(const float* x, int64_t ldx, const float* gamma, const float* beta, int64_t mod_frame_stride,
                 int rows_per_frame, int gamma_plus_one, void* out, int64_t ldo, int out_f32, int64_t rows, int D,
                 float eps, void* stream);

/* mod[s][n][i][0:dim] = scale, [dim:2dim] = shift of AdaLayerNorm i (norm1..3) in block n for sigma[s]:
 * PositionalEmbedding + map_layer0/1 + SiLU (model/models_radar_generation.py:27-33, 217-219) followed by all
 * depth*3 AdaLayerNorm.linear layers (:128-129). freqs = the fp32 frequency vector of :28-30 (length `half`).
 * ada_w: fp32 [depth*3*2*dim, dim] rows ordered (block, norm index, out feature); t_emb_ws: scratch [S, dim]. */
int rald_dit_mod_table(const float* sigma, int S, const float* freqs, int half, const float* map0_w,
                       const float* map0_b, const float* map1_w, const float* map1_b, const float* ada_w,
                       const float* ada_b, int depth, int dim, float* t_emb_ws, float* mod, void* stream);

/* Fused "evaluation boundary" on [T, C] latent rows (C <= 32, dim 512):
 *   F = proj_out(LayerNorm(h)) (model/models_radar_generation.py:230-232); D = c_skip x + c_out F (:422-429);
 *   mode 0: x_out = D                                   (EDMPrecond.forward)
 *   mode 1: d = (x - D)/sigma; x_out = x + (sigma_other - sigma) d; d_buf = d          (Euler, :265-266)
 *   mode 2: d' = (x - D)/sigma; x_out = x_base + (sigma - sigma_other)(d_buf/2 + d'/2)  (Heun, :272-273)
 *   mode 3: x_out = x * sigma (:252);  mode 4: nothing (projection only)
 *   then, if h_next != NULL: h_next = proj_in(c_in(sigma_next) * x_out) (:221, :427) for the next evaluation.
 * w_out_t: fp32 [dim][32] (proj_out transposed, zero padded); w_in_t: fp32 [C][dim]. T and rows-per-frame are
 * multiples of 16 (the row tile). pack: rald_dit_boundary_pack's image of the four weight arrays, or NULL (the raw
 * weights are then split and packed by an extra launch in front of every call). The projections multiply split-bf16
 * operands (hi + lo halves, 16+ mantissa bits) with fp32 accumulation; the update arithmetic is fp32. */
int rald_dit_boundary(const float* h, const float* ln_w, const float* ln_b, const float* w_out_t,
                      const float* w_in_t, const float* x_in, const float* x_base, float* d_buf, float* x_out,
                      float* h_next, const float* sigma, int64_t sigma_stride, const float* sigma_other,
                      int64_t sigma_other_stride, int mode, int rows_per_frame, int C, int64_t T, int dim,
                      float sigma_data, const void* pack, void* stream);

/* The weights of rald_dit_boundary split into bf16 hi + lo halves, packed and laid out in the order the kernel's tensor-core
 * fragments read them, plus the column sums the algebraic LayerNorm needs: rald_dit_boundary_pack_bytes() bytes at a
 * 16-byte aligned `pack`. Made once per weight version by the runtimes (rald_dit_weights.boundary_pack). */
int64_t rald_dit_boundary_pack_bytes(void);
int rald_dit_boundary_pack(const float* ln_w, const float* ln_b, const float* w_out_t, const float* w_in_t, int C,
                           void* pack, void* stream);

/* Radar conditioning tokens: Linear(cz -> dim) of the encoder output [B, nr, na, ne, cz] (channels last) plus the
 * range / azimuth / elevation embeddings (model/models_radar_generation.py:390-405). Either output may be NULL. */
int rald_radar_tokens(const float* feat, int B, int nr, int na, int ne, int cz, const float* w, const float* b,
                      const float* r_emb, const float* a_emb, const float* e_emb, int dim, float* tok_f32,
                      void* tok_bf16, void* stream);

/* ---- denoiser runtime: packed weights + workspace + whole-evaluation / whole-sampler entry points ---- */
typedef struct rald_dit_weights {
  int32_t depth, dim, heads, channels, n_latents, ctx_len;
  float sigma_data;
  int32_t precise;     /* 0: bf16 weights (default). 1: every weight matrix below is a split pair [rows][2*cols] =
                        * [W_hi | W_lo] (rald_gemm_bf16_wsplit: 16 mantissa bits of the fp32 weight, twice the GEMM work)
                        * and the GEGLU uses the erf GELU; the fused cross-attention operands must not be supplied */
  /* bf16, stacked over depth */
  const void* w_qkv;   /* [depth][3*dim][dim]  rows: to_q | to_k | to_v of attn1 */
  const void* w_o1;    /* [depth][dim][dim]    attn1.to_out.0.weight */
  const void* w_q2;    /* [depth][dim][dim]    attn2.to_q.weight */
  const void* w_o2;    /* [depth][dim][dim]    attn2.to_out.0.weight */
  const void* w_ff1;   /* [depth][8*dim][dim]  ff.net.0.proj.weight, GEGLU-packed (16 value + 16 gate rows) */
  const void* w_ff2;   /* [depth][dim][4*dim]  ff.net.2.weight */
  /* fp32 */
  const float* b_o1;   /* [depth][dim] */
  const float* b_o2;   /* [depth][dim] */
  const float* b_ff1;  /* [depth][8*dim] GEGLU-packed */
  const float* b_ff2;  /* [depth][dim] */
  const float* ln_w;   /* [dim] model.norm.weight */
  const float* ln_b;   /* [dim] */
  const float* proj_in_t;   /* [channels][dim]  proj_in.weight transposed */
  const float* proj_out_t;  /* [dim][32]        proj_out.weight transposed, zero padded to 32 */
  const void* boundary_pack; /* optional: rald_dit_boundary_pack's image of (ln_w, ln_b, proj_out_t, proj_in_t), made once
                              * per weight version; NULL = the evaluation boundary packs the raw weights on every launch */
} rald_dit_weights;

typedef struct rald_dit_workspace {
  int32_t max_frames;  /* micro-batch size the buffers below are sized for; T = max_frames * n_latents */
  int32_t xattn_split_below;  /* with xattn_kp set: micro-batches of fewer frames than this run the folded cross-attention
                               * as two GEMMs (rald_xattn_split) instead of the fused kernel — same result bit for bit,
                               * fills the machine at small batch; 0 = always the fused kernel */
  float* h;      /* [T][dim]   fp32 residual stream */
  void* xn;      /* [T][dim]   bf16 normalised operand */
  void* qkv;     /* [T][3*dim] bf16 */
  void* att;     /* [T][dim]   bf16 */
  void* ff;      /* [T][4*dim] bf16 */
  float* x_tmp;  /* [T][channels] */
  float* d_tmp;  /* [T][channels] */
  /* optional (NULL = unfused attn2 path): context operands of the fused cross-attention, built by rald_xattn_fold
   * for xattn_frames frames; frame f of a call to rald_dit_forward / rald_dit_sample uses entry xattn_frame0 + f of
   * them (xattn_frame0 > 0: the call processes a sub-batch of the sample, e.g. one of several concurrent chains) */
  const void* xattn_kp;   /* bf16 [depth][8][xattn_frames][64][dim] */
  const void* xattn_vt;   /* fp16 [depth][8][dim][xattn_frames*64] */
  int32_t xattn_frames;
  int32_t xattn_frame0;
} rald_dit_workspace;

/* One EDMPrecond.forward (model/models_radar_generation.py:412-430) for `frames` frames given precomputed
 * conditioning: mod = adaLN table rows for each frame's sigma ([frames or 1][depth][3][2*dim], frame stride
 * mod_frame_stride, 0 = shared), ctxkv = bf16 [frames*ctx_len][depth*2*dim] (per block: K | V projections of
 * the conditioning tokens, attn2.to_k / to_v). x, out: fp32 [frames][n_latents][channels]. */
int rald_dit_forward(const rald_dit_weights* w, const rald_dit_workspace* ws, const float* x, const float* sigma,
                     int64_t sigma_stride, const float* mod, int64_t mod_frame_stride, const void* ctxkv,
                     float* out, int frames, void* stream);

/* edm_sampler (model/models_radar_generation.py:235-275, S_churn = 0): sigmas = device fp32 [num_steps + 1]
 * (last = 0), mod = [num_steps][depth][3][2*dim], latents = unit normal fp32 [frames][n_latents][channels].
 * x_out receives the final latents; trace (optional) [num_steps][frames][n_latents][channels] receives x_next
 * after every step. */
int rald_dit_sample(const rald_dit_weights* w, const rald_dit_workspace* ws, const float* latents,
                    const float* sigmas, int num_steps, const float* mod, const void* ctxkv, float* x_out,
                    float* trace, int frames, void* stream);

/* Fused attn2 sub-layer (model/models_radar_generation.py:35-76 called at :167) for 8 heads x 64 and a 64-token
 * context that is constant over the sampler's network evaluations. rald_xattn_fold (once per sample) folds
 * attn2.to_q into the keys and attn2.to_out into the values of every block:
 *   ctxkv_bf16 [frames*64][depth*2*dim] (per block K | V, BOTH bf16), wq_t_scaled bf16 [depth][dim in][dim q] =
 *   attn2.to_q.weight transposed times log2(e)/sqrt(64), w_o bf16 [depth][dim][dim] = attn2.to_out.0.weight
 *   -> kp bf16 [depth][8][frames][64][dim], vt fp16 [depth][8][dim][frames*64].
 * rald_xattn_fused then computes, for ONE block's kp / vt slices, h[T][dim] += softmax_per_head(xn kp^T) vt^T + bias
 * in a single kernel (xn bf16 [T][dim] = adaLN2(h); T = frames * rows_per_frame; frame f of the call reads entry
 * frame0 + f of operands built for total_frames frames). */
int rald_xattn_fold(const void* ctxkv_bf16, const void* wq_t_scaled, const void* w_o, int depth, int frames, void* kp,
                    void* vt, void* stream);
int rald_xattn_fused(const void* xn, const void* kp, const void* vt, const float* bias, float* h, int frames,
                     int rows_per_frame, int frame0, int total_frames, void* stream);
/* The same sub-layer as TWO tcgen05 GEMMs for batches too small to fill the machine with 128-row fused tiles (4 tiles
 * per frame): probs (fp16 scratch [T][512]) = per-head softmax(xn kp^T) in the first GEMM's epilogue, then
 * h += probs vt^T + bias with a TMA reduce-add epilogue. Same operands and arithmetic order as rald_xattn_fused: the
 * results are bit-identical. */
int rald_xattn_split(const void* xn, const void* kp, const void* vt, const float* bias, float* h, void* probs_f16,
                     int frames, int rows_per_frame, int frame0, int total_frames, void* stream);
/* Debug hook like rald_gemm_debug_buffer for rald_ae_query (phase list in csrc/ae_query.cu). */
int rald_ae_query_debug_buffer(unsigned long long* dev_buf);
/* Debug hook like rald_gemm_debug_buffer: CTA 0 of the fused kernel stores %globaltimer stamps of its first 4 tiles at
 * dev_buf[tile*16 + i] (phase list in csrc/xattn.cu). */
int rald_xattn_debug_buffer(unsigned long long* dev_buf);

/* ---- VecSet autoencoder (model/models_ae.py) ---- */
typedef struct rald_ae_weights {
  int32_t depth, dim, heads, latent_dim, n_latents;
  int32_t precise;     /* 0: bf16 weights, logistic-form GELU. 1: every weight matrix below is a split pair
                        * [rows][2*cols] = [W_hi | W_lo] (rald_gemm_bf16_wsplit) and the GEGLU uses the erf GELU */
  /* bf16, stacked over depth */
  const void* w_qkv;   /* [depth][3*dim][dim]  layers.N.0.fn: to_q | to_kv (k rows, then v rows) */
  const void* w_o;     /* [depth][dim][dim]    layers.N.0.fn.to_out.weight */
  const void* w_ff1;   /* [depth][8*dim][dim]  layers.N.1.fn.net.0.weight, GEGLU-packed */
  const void* w_ff2;   /* [depth][dim][4*dim]  layers.N.1.fn.net.2.weight */
  /* fp32 */
  const float* b_o;    /* [depth][dim] */
  const float* b_ff1;  /* [depth][8*dim] GEGLU-packed */
  const float* b_ff2;  /* [depth][dim] */
  const float* ln1_w;  /* [depth][dim] layers.N.0.norm */
  const float* ln1_b;
  const float* ln2_w;  /* [depth][dim] layers.N.1.norm */
  const float* ln2_b;
  const float* proj_wt; /* [latent_dim][dim] proj.weight transposed; NULL (with latent_dim == dim): deterministic
                         * AutoEncoder.decode (models_ae.py:260-264), the latents are the residual stream */
  const float* proj_b;  /* [dim] */
} rald_ae_weights;

/* x_out[frames*n_latents][dim] (fp32) = latent stack of KLAutoEncoder.decode: proj followed by depth x
 * (self-attention, GEGLU feed-forward) with pre-LayerNorm and residuals (model/models_ae.py:410-414).
 * z: fp32 [frames][n_latents][latent_dim]. Uses the same workspace layout as the denoiser. */
int rald_ae_stack(const rald_ae_weights* w, const rald_dit_workspace* ws, const float* z, float* x_out,
                  int frames, void* stream);

/* out[T][512] = x[T][K] @ wt[K][512] + b in fp32 (K <= 64): KLAutoEncoder.proj (model/models_ae.py:346, 410). */
int rald_linear_smallk(const float* x, int K, const float* wt, const float* b, float* out, int64_t T, int N,
                       void* stream);

/* out[row] = (LayerNorm(x[row]) * g + b) . w over 512-wide fp32 rows: the folded value path
 * v' = LN_ctx(x) (W_v^T W_out^T w_o^T) of the decoder cross-attention (model/models_ae.py:89, 103-105, 424). */
int rald_ln_dot_rows(const float* x, const float* g, const float* b, const float* w, float* out, int64_t rows,
                     int D, float eps, void* stream);

/* Decoder queries (model/models_ae.py:417-424), folded form: logits[b][q] = softmax(LN(point_embed(q)) K'_b^T /
 * sqrt(dim)) . v'_b + c0_b. queries fp32 [B][Q][3]; wpe_bf16 [512][64] = point_embed.mlp.weight zero-padded
 * from 51 to 64 inputs; pe_bias = point_embed.mlp.bias; ln_g/ln_b = decoder_cross_attn.norm; kprime_bf16
 * [B*512][512]; vprime [B][512]; c0 [B]; freq24_host = HOST pointer to the 24 non-zero entries of
 * point_embed.basis (x, y, z blocks of 8). */
int rald_ae_query(const float* queries, int B, int64_t Q, const void* wpe_bf16, const float* pe_bias,
                  const float* ln_g, const float* ln_b, const void* kprime_bf16, const float* vprime,
                  const float* c0, const float* freq24_host, float* logits, int dim, int n_latents, void* stream);

/* ---- VecSet autoencoder, encode side (model/models_ae.py:351-405) ---- */

/* feat bf16 [rows, 64] = [sin(p f) (24), cos(p f) (24), p (3), 0 (13)] of PointEmbed (:128-137) for pts fp32
 * [B, N, 3]; with idx (int64 [B, M], indices into each cloud) the rows are the gathered points pts[b][idx[b][i]]
 * (also copied to `gathered` [B, M, 3] when given). freq24_host: HOST pointer, see rald_ae_query. */
int rald_point_features(const float* pts, const int64_t* idx, int B, int64_t N, int64_t M, const float* freq24_host,
                        void* feat_bf16, float* gathered, void* stream);

/* Farthest point sampling of M of the N points of each of B clouds (pts fp32 [B, N, 3]) -> out_idx int64 [B, M],
 * indices INTO EACH CLOUD in selection order. Replaces torch_cluster.fps (model/models_ae.py:243, 368; the global
 * index of the reference is b*N + out_idx[b][i]). Deterministic: start index 0, squared distance
 * (dx*dx + dy*dy) + dz*dz in individually rounded fp32 operations, ties -> lowest index. N <= 16384. */
int rald_fps(const float* pts, int B, int N, int M, int64_t* out_idx, void* stream);

/* DiagonalGaussianDistribution (:141-163): ml fp32 [B*rows, ld] = (mean | logvar) columns [0,C) and [C,2C); writes
 * mean, clamp(logvar, -30, 20), z = mean + exp(logvar/2) noise (when z != NULL) and kl[b]. */
int rald_ae_posterior(const float* ml, int64_t ld, const float* noise, int B, int rows_per_frame, int C, float* mean,
                      float* logvar, float* z, float* kl, void* stream);

typedef struct rald_ae_enc_weights {
  int32_t dim, n_latents, latent_dim, heads;
  int32_t query_type;            /* 0 = point (FPS), 1 = learnable, 2 = mix */
  int32_t stats_rows;            /* 2*latent_dim rounded up to a multiple of 32 */
  float freq24[24];              /* non-zero entries of point_embed.basis */
  const void* wpe;               /* bf16 [dim][64] point_embed.mlp.weight zero padded */
  const float* pe_bias;
  const void* mix_q;             /* bf16 [n_latents][dim] = to_q(LN(d_latents)) (batch independent) */
  const void* mix_wkv;           /* bf16 [2*dim][dim] mix_attn_layer.fn.to_kv.weight (k rows, then v rows) */
  const void* mix_wo;            /* bf16 [dim][dim] */
  const float* mix_bo;
  const float* s_latents;        /* fp32 [n_latents][dim] */
  const void* wproj;             /* bf16 [dim][dim] query_proj */
  const float* bproj;
  const float* latents;          /* fp32 [n_latents][dim] (learnable queries) */
  const float* ca_ln_w; const float* ca_ln_b;      /* cross_attend_blocks.0.norm */
  const float* ca_lnc_w; const float* ca_lnc_b;    /* cross_attend_blocks.0.norm_context */
  const void* ca_wq; const void* ca_wkv; const void* ca_wo; const float* ca_bo;
  const float* ff_ln_w; const float* ff_ln_b;      /* cross_attend_blocks.1 */
  const void* ff_w1; const float* ff_b1; const void* ff_w2; const float* ff_b2;   /* GEGLU-packed w1/b1 */
  const void* w_stats;           /* bf16 [stats_rows][dim]: mean_fc rows, logvar_fc rows, zero rows; NULL: deterministic
                                  * AutoEncoder.encode (models_ae.py:226-257), ml_out receives x itself [B*n_latents][dim] */
  const float* b_stats;
} rald_ae_enc_weights;

typedef struct rald_ae_enc_workspace {
  int32_t max_points, _pad;      /* buffers below are sized for one frame of max_points (n_pad = ceil32) */
  void* feat;                    /* bf16 [max_points][64] */
  float* pe32;                   /* fp32 [max_points][dim] */
  void* pe16;                    /* bf16 [max_points][dim] */
  void* kbuf;                    /* bf16 [max_points][dim] */
  void* vt;                      /* bf16 [dim][n_pad] */
  float* scores;                 /* fp32 [n_latents][n_pad] */
  void* prob;                    /* bf16 [n_latents][n_pad] */
  float* x;                      /* fp32 [n_latents][dim] */
  void* xq; void* att; void* xn; /* bf16 [n_latents][dim] */
  void* ff;                      /* bf16 [n_latents][4*dim] */
} rald_ae_enc_workspace;

/* KLAutoEncoder.encode up to the posterior parameters (:351-399): pc fp32 [B, N, 3] -> ml_out fp32
 * [B*n_latents][stats_rows] = (mean | logvar | padding). fps_idx (int64 [B, n_latents]) receives the sampled indices
 * for query_type 0 and may be NULL otherwise. */
int rald_ae_encode_stats(const rald_ae_enc_weights* w, const rald_ae_enc_workspace* ws, const float* pc, int B, int N,
                         int64_t* fps_idx, float* ml_out, void* stream);

/* ---- radar-cube encoder (model/models_radar_encoder.py) ---- */

/* out[B, D/s, H/s, W/s, Cout] (fp32, channels last) = Conv3d 3x3x3 of x[B, D, H, W, Cin] (bf16, channels last) on
 * tcgen05 as an implicit GEMM; stride 1 pads 1 on both sides (ResnetBlock convs :63-72, conv_out :208-214), stride 2
 * pads 1 on the HIGH side only (Downsample :34-41). w_packed: bf16 [w_rows][27*Cin] with K index = tap*Cin + ci,
 * tap = kd*9 + kh*3 + kw, w_rows = Cout rounded up to the N tile (32 / 64 / 128) with zero rows; bias padded
 * likewise. resid (optional, fp32, same shape as out, may alias out) is added in the epilogue. Cin % 64 == 0. */
int rald_conv3d_cl(const void* x_bf16, const void* w_packed, int w_rows, const float* bias, const float* resid,
                   float* out, int B, int D, int H, int W, int Cin, int Cout, int stride, void* stream);

/* conv_in (:161-163): x fp32 [B, D, H, W, Cin<=4] -> out fp32 [B, D, H, W, Cout]; w = PyTorch layout
 * [Cout][Cin][3][3][3] fp32. */
int rald_enc_conv_in(const float* x, const float* w, const float* bias, float* out, int B, int D, int H, int W,
                     int Cin, int Cout, void* stream);

/* GroupNorm(groups, eps) statistics of x fp32 [B, V, C]: stats[b][g] = {sum, sum of squares} in double. */
int rald_gn_stats(const float* x, int B, int64_t V, int C, int groups, double* stats, void* stream);

/* out bf16 [B, V, C] = mode 0: swish(GN(x)) (Normalize + nonlinearity :5-12); 1: GN(x); 2: x (cast only). */
int rald_gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* out_bf16, int B,
                  int64_t V, int C, int groups, float eps, int mode, void* stream);

/* AttnBlock core (:121-133): qkv fp32 [B*n, 3C] (q | k | v), one head of width C over the n <= 64 voxels of each
 * frame; out bf16 [B*n, C] = softmax(q k^T C^-0.5) v. */
int rald_enc_attn(const float* qkv, void* out_bf16, int B, int n, int C, void* stream);

typedef struct rald_enc_conv {   /* 3x3x3 conv (packed as for rald_conv3d_cl) or 1x1x1 conv (bf16 [cout][cin]) */
  const void* w;                 /* NULL = layer absent */
  const float* b;
  int32_t cin, cout, w_rows, _pad;
} rald_enc_conv;
typedef struct rald_enc_norm { const float* g; const float* b; } rald_enc_norm;
typedef struct rald_enc_resblock { rald_enc_norm n1; rald_enc_conv c1; rald_enc_norm n2; rald_enc_conv c2;
                                   rald_enc_conv nin; } rald_enc_resblock;
typedef struct rald_enc_attnblock { rald_enc_norm n; rald_enc_conv qkv; rald_enc_conv proj; } rald_enc_attnblock;
#define RALD_ENC_MAX_LEVELS 8
#define RALD_ENC_MAX_BLOCKS 4
typedef struct rald_enc_level {
  int32_t n_blocks, n_attn;      /* n_attn is 0 or n_blocks */
  rald_enc_resblock block[RALD_ENC_MAX_BLOCKS];
  rald_enc_attnblock attn[RALD_ENC_MAX_BLOCKS];
  rald_enc_conv down;            /* stride-2 conv, absent on the last level */
} rald_enc_level;
typedef struct rald_enc_weights {
  int32_t n_levels, in_ch, ch, z_ch, groups, _pad;
  float eps;
  int32_t _pad2;
  const float* conv_in_w;        /* fp32 [ch][in_ch][3][3][3] */
  const float* conv_in_b;
  rald_enc_level level[RALD_ENC_MAX_LEVELS];
  rald_enc_resblock mid1, mid2;
  rald_enc_attnblock mid_attn;
  rald_enc_norm norm_out;
  rald_enc_conv conv_out;
} rald_enc_weights;
typedef struct rald_enc_workspace {
  int32_t max_frames, _pad;
  int64_t elems;                 /* capacity of each buffer below in elements (>= max_frames * D*H*W * widest C) */
  float* x;                      /* fp32 residual stream */
  float* y;                      /* fp32 second stream buffer (shortcut / downsample output) */
  float* t;                      /* fp32 conv1 output, attention qkv */
  void* xb;                      /* bf16 normalised operand */
  double* stats;                 /* [max_frames][groups][2] */
} rald_enc_workspace;

/* Encoder.forward (model/models_radar_encoder.py:216-241) for channels-last input x fp32 [B, D, H, W, in_ch]:
 * out fp32 [B, D/2^(L-1), H/2^(L-1), W/2^(L-1), z_ch]. Frames are processed in micro-batches of ws->max_frames. */
int rald_radar_encoder(const rald_enc_weights* w, const rald_enc_workspace* ws, const float* x, float* out, int B,
                       int D, int H, int W, void* stream);

/* ---- post-processing of decoded occupancy (engine_generation.py:283-289, 313-315) ---- */

/* Stable stream compaction of the occupied queries of each frame: for every q (in query order) with
 * logits[b][q] > threshold, points[b][n] = queries[b][q] * scale + offset (inverse_norm_points, utils/utils.py:50-76;
 * scale_offset_host = HOST pointer to {sx, sy, sz, ox, oy, oz} or NULL for identity), optionally followed by
 * polar2cartesian (dataset_preprocessor/lidar.py:57-63); index[b][n] = q (optional). counts[b] = number of occupied
 * queries (may exceed cap; only the first cap are stored). block_ws: int32 scratch of rald_occupancy_ws_elems(B, Q). */
int rald_occupancy_compact(const float* logits, const float* queries, int B, int64_t Q, float threshold,
                           const float* scale_offset_host, int polar2cart, int64_t cap, float* points, int32_t* index,
                           int32_t* counts, int32_t* block_ws, void* stream);
int64_t rald_occupancy_ws_elems(int B, int64_t Q);

/* ---- the eval loop either side of sampler + decoder ---- */

/* Second-pass ("refine_query") query set of one frame, engine_generation.py:291-297:
 *   aug_query_helper(pred, aug_num, pc_range, voxel_size, aug_scale)  datasets/utils/query_helper.py:3-43
 *   norm_points(., pc_range, ...)                                      utils/utils.py:78-104
 * points f32 [cap, 3] = the first pass's occupied points (inverse-normalised, as written by
 * rald_occupancy_compact), *count (device int32) of them valid. Rows [0, N) of out f32 [aug_num, 3] are the points
 * themselves, rows N + g are float32(clip(points[sel[g]] + (u[g]*2-1) * (voxel_size * scales[g]), range)) (fp64
 * arithmetic as numpy's), all normalised back with (p - offset) / scale in fp32 (norm_scale_offset_host =
 * {sx, sy, sz, ox, oy, oz}). The three draw arrays (sel i32 [aug_num], scales i32 [aug_num] in [1, aug_scale],
 * uniforms f64 [aug_num, 3]; indexed by g) are either all given — numpy-parity mode: the host draws them with
 * np.random in the reference's order — or all NULL: Philox4x32-10 on the device, keyed (seed, g). */
int rald_refine_queries(const float* points, const int32_t* count, int64_t cap, int64_t aug_num, const int32_t* sel,
                        const int32_t* scales, const double* uniforms, uint64_t seed, int aug_scale,
                        const double* voxel_size_host, const double* pc_range_host, const float* norm_scale_offset_host,
                        float* out, void* stream);

/* cal_metrics / chamfer_distance (utils/utils.py:116-142) for B frames at once: pred f32 [B, pred_cap, 3] with
 * pred_counts[b] valid rows, gt f32 [B, gt_cap, 3] with gt_counts[b] valid rows (gt_counts NULL: gt_fixed rows in
 * every frame). out f64 [B, 3] = {0.5*mean_gt(NN dist to pred) + 0.5*mean_pred(NN dist to gt), mean pred->gt,
 * mean gt->pred}; +inf where a side is empty. Brute-force fp32 nearest-neighbour search, the winning distance is
 * recomputed in fp64 (the reference's cKDTree works in fp64). ws: f64 scratch of rald_chamfer_ws_elems(). */
int rald_chamfer(const float* pred, const int32_t* pred_counts, int64_t pred_cap, const float* gt,
                 const int32_t* gt_counts, int64_t gt_cap, int gt_fixed, int B, double* out, double* ws, void* stream);
int64_t rald_chamfer_ws_elems(int B, int64_t pred_cap, int64_t gt_cap);

/* Coloradar_dataset.process_radar_data (datasets/aligned_coloradar/Coloradar_dataset.py:432-475): raw f32
 * [B, R, A, E, C] (C >= 2: intensity dB, doppler, ..., valid mask last) -> out f32 [B, R, A_up, E_up, channels_out]:
 * channel 0 = clip(I, 0, max_intensity) / max_intensity (zeros if !norm_intensity), channel 1 = doppler * mask
 * (/ max_dopp if norm_dopp), then bilinear upsampling of the (A, E) plane with align_corners=True (A_up = A and
 * E_up = E: none). channels_out = 1 writes channel 0 only — the tensor the radar encoder consumes
 * (model/models_radar_generation.py:378). */
int rald_radar_cube_prep(const float* raw, int B, int R, int A, int E, int C, int A_up, int E_up, int channels_out,
                         int norm_intensity, float max_intensity, int norm_dopp, float max_dopp, float* out,
                         void* stream);

/* ---- training-loop helpers that need no autograd ---- */

/* update_ema (engine_generation.py:29-39) over a whole parameter list in ONE launch:
 *   target[i] = fma(source[i], one_minus_rate, rn(target[i] * rate))   -- ATen's mul_(rate).add_(src, alpha=1-rate)
 * table_dev: device int64 [4 * n_tensors + 1] = { target pointers [n] | source pointers [n] | element counts [n] |
 * first chunk of each tensor [n + 1] (exclusive prefix sum of ceil(count / rald_ema_chunk_elems()), last entry =
 * total_chunks) }. All tensors fp32, contiguous; target and source of a pair must not overlap. total_elems is used
 * for launch accounting only (12 B of HBM traffic per element). */
int rald_ema_update(const int64_t* table_dev, int n_tensors, int64_t total_chunks, int64_t total_elems, float rate,
                    float one_minus_rate, void* stream);
int rald_ema_chunk_elems(void);

/* Occupancy accuracy / IoU of the AE evaluation loops (engine_generation.py:376-385 in cache_latents; engine_ae.py's
 * evaluate has the same lines): pred = logits >= threshold. logits, labels f32 [B, Q] (labels 0 / 1);
 * out f32 [B, 2] = { mean(pred == labels), sum(pred * labels) / count(pred + labels > 0) + 1e-5 } per frame (the
 * reference then averages over the batch); ws: int32 [3 * B] scratch. */
int rald_occupancy_iou(const float* logits, const float* labels, int B, int64_t Q, float threshold, float* out,
                       int32_t* ws, void* stream);

/* ---- the backward pass of the denoiser (EDMLoss under autograd,
 * model/models_radar_generation.py:277-295 called at engine_generation.py:89-110). The matrix products of the
 * backward pass are calls of rald_gemm_bf16: dgrad = dY W with a transposed bf16 copy of the weight as the W operand,
 * wgrad = dY^T X over K = rows with transposed activations and out_mode 1 / resid = out (fp32 accumulation by the
 * TMA reduce-add epilogue). The entry points below are what surrounds them. ---- */

/* rald_attn_d64 that also writes the softmax statistics stats f32 [frames*Sq][heads][2] = (row maximum m in log2
 * units with the scale folded in, row sum l of 2^(s - m)): what rald_attn_d64_bwd recomputes the probabilities from. */
int rald_attn_d64_stats(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                        int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, float* stats, void* stream);

/* Backward of rald_attn_d64 (autograd through the einsum / softmax / einsum of CrossAttention.forward,
 * model/models_radar_generation.py:66-75): given Q, K (bf16), V_centred (bf16: the forward's V minus its mean over the
 * keys of each frame, per column — rald_center_cast_f16_bf16; dS is invariant under that shift and the products of the
 * backward pass are far better conditioned with it, see csrc/attn_bwd.cu), the forward's statistics and dO (bf16), writes
 * dQ [frames*Sq][..], dK and dV [frames*Skv][..] as bf16 into columns [h*64, h*64+64) of rows of pitch lddq / lddk /
 * lddv. Sq % 128 == 0; Skv = 64 or a multiple of 128; heads <= 8. Scratch: lse2_ws, dsum_ws f32 [frames*heads*Sq] each.
 * Deterministic (no atomics). */
int rald_attn_d64_bwd(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V_centred, int64_t ldv,
                      const void* dO, int64_t lddo, const float* stats, float* lse2_ws, float* dsum_ws, void* dQ,
                      int64_t lddq, void* dK, int64_t lddk, void* dV, int64_t lddv, int frames, int heads, int Sq, int Skv,
                      float scale, void* stream);

/* in [R, C] (f32 when in_f32, else bf16; pitch ld_in) -> out_bf16 [R, C] (optional) and out_t_bf16 [C, R] (optional,
 * pitch ld_t >= R): the bf16 / transposed operands of the dgrad and wgrad GEMMs. colsum_partial (optional; needs R and
 * C multiples of 64 and 16-byte aligned rows): f32 [R/64][C], the column sums of every 64-row tile of the INPUT —
 * finished by rald_colsum_finish(partial, R/64, C, out, accumulate) into the bias gradient, fixed summation order. */
int rald_cast_transpose(const void* in, int in_f32, int64_t ld_in, int64_t R, int64_t C, void* out_bf16, int64_t ld_out,
                        void* out_t_bf16, int64_t ld_t, float* colsum_partial, void* stream);
int rald_colsum_finish(const float* partial, int chunks, int64_t C, float* out, int accumulate, void* stream);

/* out_bf16[f][j][c] = bf16(in_f16[f][j][c] - mean_j in_f16[f][j][c]) over frames x rows_per_frame rows of C columns
 * (pitches in elements): the fp16 V columns rald_attn_d64 consumed, centred and re-encoded for rald_attn_d64_bwd. */
int rald_center_cast_f16_bf16(const void* in_f16, int64_t ld_in, void* out_bf16, int64_t ld_out, int frames,
                              int rows_per_frame, int64_t C, void* stream);

/* out[c] (+)= sum over rows of in[r][c] (bias gradients); two deterministic stages through partial_ws
 * (f32, >= min(512, ceil(R/256)) * C elements). */
int rald_colsum(const void* in, int in_f32, int64_t ld, int64_t R, int64_t C, float* partial_ws, int64_t ws_elems,
                float* out, int accumulate, void* stream);

/* Backward of rald_ln_rows (AdaLayerNorm.forward :127-131 / nn.LayerNorm): x f32 [rows][512] (the forward input),
 * dy bf16 [rows][512], gamma as for rald_ln_rows. dh f32 [rows][512] (+)= dx. dparam[g][0][:] (+)= sum dy * xhat,
 * dparam[g][1][:] (+)= sum dy over the rows of group g (= frame when rows_per_frame > 0, else all rows), group pitch
 * dparam_group_stride and second-vector offset dparam_which_stride (in floats). rows and rows_per_frame multiples
 * of 64. partial_ws: f32 scratch of rows / 64 * 1024 elements. */
int rald_ln_bwd(const float* x, const void* dy_bf16, const float* gamma, int64_t mod_frame_stride, int rows_per_frame,
                int gamma_plus_one, float* dh, int accumulate_dh, float* partial_ws, int64_t ws_elems, float* dparam,
                int64_t dparam_group_stride, int64_t dparam_which_stride, int accumulate_dparam, int64_t rows, int D,
                float eps, void* stream);

/* GEGLU (:88-95) on a materialised projection u bf16 [T][2*inner] (value columns, then gate columns), erf GELU:
 * g bf16 [T][inner] = value * gelu(gate); backward: du = [dg * gelu(gate) | dg * value * gelu'(gate)]. */
int rald_geglu_fwd(const void* u_bf16, int64_t T, int inner, void* g_bf16, void* stream);
int rald_geglu_bwd(const void* u_bf16, const void* dg_bf16, int64_t T, int inner, void* du_bf16, void* stream);

/* C[M, N] = alpha * op(A) op(B) + beta * C, fp32 row-major, op = transpose when trans_* != 0 (small operands only:
 * the timestep-embedding MLP :217-219 and its gradients). */
int rald_sgemm_f32(int trans_a, int trans_b, int M, int N, int K, float alpha, const float* A, int64_t lda, const float* B,
                   int64_t ldb, float beta, float* C, int64_t ldc, void* stream);

/* Gradients of the token projection and position embeddings of process_radar_cond (:390-405) from dtok f32
 * [B*nr*na*ne][dim] and the encoder features feat f32 [B*nr*na*ne][cz]: dw [dim][cz], db [dim], dr_emb [nr][dim],
 * da_emb [na][dim], de_emb [ne][dim] (rows of the embedding tables beyond nr / na / ne receive no gradient). */
int rald_radar_tokens_bwd(const float* dtok, const float* feat, int B, int nr, int na, int ne, int cz, int dim, float* dw,
                          float* db, float* dr_emb, float* da_emb, float* de_emb, void* stream);

/* ---- backward of the radar-cube encoder (training with unfreeze_radar_enc: true; autograd through
 * model/models_radar_encoder.py). Convolution dgrad = rald_conv3d_cl on flipped / transposed weights, wgrad =
 * rald_gemm_bf16_accum_shift per tap, 1x1 convolutions = GEMMs; the rest: ---- */

/* GroupNorm(groups, eps) (+ swish when swish != 0) backward (Normalize + nonlinearity :5-12): x, dy fp32 [B, V, C], stats
 * = rald_gn_stats of x. dx = d loss / d x (+ add when given, e.g. the identity-shortcut gradient); sums f64 [B][C][2] =
 * {sum_v dyh * xhat, sum_v dyh} (d gamma / d beta = their sums over B). */
int rald_gn_bwd(const float* x, const float* dy, const double* stats, const float* gamma, const float* beta, int B, int64_t V,
                int C, int groups, float eps, int swish, double* sums, const float* add, float* dx, void* stream);

/* in [B, D, H, W, C] (f32 or bf16, channels last) -> out_t bf16 [copies * copy_rows][ld]: element (c, P) of copy r at
 * row r * copy_rows + c, column P + pos_bias - r, P = flattened index of voxel (dil*d+1, dil*h+1, dil*w+1) on the
 * zero-padded grid [B][dil*D+2][dil*H+2][Wp] (Wp % 8 == 0, >= dil*W+2). out_t must be zero-initialised, ld >= grid + 8.
 * dil = 2 places a stride-2 convolution's output gradient on its input grid; copies = 3 writes the three kw-shifted
 * copies of a convolution input (TMA box origins must be 16-byte aligned: only the kd / kh part of a tap offset can be
 * an operand shift of rald_gemm_bf16_accum_shift). colsum (optional, f64 [C], f32 input only) receives the column sums of
 * the input over all voxels: the convolution's bias gradient from the pass that already reads dY. panel_len > 0: the
 * K-panel-major layout of rald_gemm_bf16_accum_taps, [n_panels][copies*copy_rows][halo | panel_len | halo] (ld unused). */
int rald_enc_pad_transpose(const void* in, int in_f32, int B, int D, int H, int W, int C, int dil, int Wp, int copies,
                           int copy_rows, int pos_bias, void* out_t_bf16, int64_t ld, int panel_len, int halo, int n_panels,
                           double* colsum, void* stream);

/* out bf16 [B, 2D, 2H, 2W, C] (zero-initialised by the caller) with out[2d+1, 2h+1, 2w+1] = in[d, h, w]: the operand of
 * the stride-2 convolution's dgrad (Downsample :34-41) as a stride-1 convolution with flipped weights. */
int rald_enc_stuff(const float* in, int B, int D, int H, int W, int C, void* out_bf16, void* stream);

/* Backward of rald_enc_attn (AttnBlock :121-133): qkv fp32 [B*n, 3C], dO fp32 [B*n, C] -> dqkv fp32 [B*n, 3C]. */
int rald_enc_attn_bwd(const float* qkv, const float* dO, float* dqkv, int B, int n, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RALD_B200_H */
