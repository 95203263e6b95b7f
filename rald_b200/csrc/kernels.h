// Internal C++ launch functions (one per kernel family). The public C ABI in api.cu forwards to these.
// Every function enqueues work on `stream`, allocates nothing, and returns 0 or a negative error code
// (message retrievable through rald_last_error()).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rald {

// gemm.cu ------------------------------------------------------------------------------------------
// out = epilogue(A[M,K] @ W[N,K]^T); A, W bf16 row-major. out_mode 0: bf16 [M,N]; 1: fp32 [M,N] (+resid);
// 2: GEGLU -> bf16 [M,N/2] with W/bias rows packed in 16-value/16-gate groups. bn_hint 0 = auto.
int gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
              const float* resid, int64_t ldr, int M, int N, int K, int out_mode, int bn_hint,
              cudaStream_t stream);

// same with the fp32 residual taken from row (row % resid_mod) of a [resid_mod, N] table when resid_mod > 0;
// the residual is also honoured by the bf16 output mode (added before rounding).
int gemm_bf16_ex(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
                 const float* resid, int64_t ldr, int64_t resid_mod, int M, int N, int K, int out_mode, int bn_hint,
                 cudaStream_t stream);

// out (fp32 [M,N], pitch ldo) += A W^T with the K extent split over several CTAs per output tile when the tiles alone do
// not fill the machine (TMA reduce-add epilogue; summation order across splits not fixed): the weight-gradient GEMMs of
// the training step, K = rows of the batch.
// w_k_off: W is read that many columns to the right of A (columns outside [0, K) read zeros).
int gemm_bf16_accum_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, int w_k_off, float* out, int64_t ldo,
                           int M, int N, int K, cudaStream_t stream);

// n_taps shifted products in one launch: out[m][t * w_rows + n] += sum_k A[m][k] W[n][k + tap_shifts[t]] (W: w_rows rows)
// panel_len > 0: K-panel-major operands, A [K/panel_len][M][panel_len], W [K/panel_len][w_rows][panel_len + 2 w_halo]
int gemm_bf16_accum_taps(const void* A, int64_t lda, const void* W, int64_t ldw, int w_rows, int n_taps,
                         const int* tap_shifts, int panel_len, int w_halo, float* out, int64_t ldo, int M, int K,
                         cudaStream_t stream);

// bf16 output in which the columns with (col % f16_period) >= f16_start are written as fp16 instead: the V
// projections feeding attn_d64 (fp16 probabilities x fp16 values on the tensor cores).
int gemm_bf16_f16cols(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo, const float* bias,
                      int M, int N, int K, int f16_start, int f16_period, cudaStream_t stream);

// SPLIT WEIGHTS: W_hilo is [N][2 K] = [W_hi | W_lo], two bf16 matrices whose sum carries 16 mantissa bits of the fp32
// weight; out = A W_hi^T + A W_lo^T accumulated in one fp32 TMEM accumulator (the A tiles are read twice). Used by the
// VecSet decoder's latent stack, where bf16 weight rounding is a common-mode error of the occupancy field
// (DESIGN.md §7). gelu_exact: erf GELU in the GEGLU epilogue. f16_period = 0: no fp16 columns.
int gemm_bf16_wsplit(const void* A, int64_t lda, const void* W_hilo, int64_t ldw, void* out, int64_t ldo,
                     const float* bias, const float* resid, int64_t ldr, int M, int N, int K, int out_mode,
                     int f16_start, int f16_period, int gelu_exact, cudaStream_t stream);

// The folded cross-attention sub-layer (xattn.cu) as TWO GEMMs for batches too small to fill the machine with its
// 128-row fused tiles (same operands, same arithmetic, bit-identical result):
//   probs[T][heads_keys] (fp16) = per-head softmax( xn[T][dim] K'^T )      K' = kp_block, bf16 [8][total_frames][64][dim]
//   h[T][dim] += probs VT^T + bias                                         VT = vt_block, fp16 [8][dim][total_frames*64]
// rows [m, m + rows_per_frame) belong to frame frame0 + m / rows_per_frame.
int gemm_xattn_scores(const void* xn, const void* kp_block, void* probs_f16, int T, int dim, int heads_keys,
                      int rows_per_frame, int frame0, int total_frames, cudaStream_t stream);
int gemm_xattn_out(const void* probs_f16, const void* vt_block, const float* bias, float* h, int T, int dim,
                   int heads_keys, int rows_per_frame, int frame0, int total_frames, cudaStream_t stream);

// While alive on this thread, every GEMM launched treats its W operand as STATIC (model weights: not written by any
// kernel in flight), which lets the kernel start streaming W before its programmatic-dependency wait.
struct GemmStaticWeights {
  GemmStaticWeights();
  ~GemmStaticWeights();
  GemmStaticWeights(const GemmStaticWeights&) = delete;
  GemmStaticWeights& operator=(const GemmStaticWeights&) = delete;
};

// attn.cu ------------------------------------------------------------------------------------------
// O[f, :, h*64:(h+1)*64] = softmax(Q K^T * scale) V per (frame f, head h); head_dim 64; Skv <= 512.
// Q rows = frames*Sq, K/V rows = frames*Skv; head h lives at columns [h*64, h*64+64) of each operand.
int attn_d64(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
             int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, cudaStream_t stream);

// one key chunk of a longer context: K / V hold kv_frame_rows >= Skv rows per frame (the pointers address the
// chunk's first row of frame 0); stats (optional) receives [frames*Sq][heads][2] = (m, l) of this chunk
int attn_d64_chunk(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                   int64_t ldo, int frames, int heads, int Sq, int Skv, int kv_frame_rows, float* stats, float scale,
                   cudaStream_t stream);
// Skv = chunks * 512 (1024 .. 4096): chunked attention + exact merge; o_chunks bf16 [chunks][frames*Sq][heads*64],
// stats fp32 [chunks][frames*Sq][heads][2] are scratch
// attn_streams.cu: the Skv = 512 case as four concurrent softmax streams per CTA (RALD_B200_ATTN_STREAMS=0: attn.cu's form)
bool attn_streams_enabled();
int attn_d64_streams(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                     int64_t ldo, int frames, int heads, int Sq, int kv_frame_rows, float* stats, float scale,
                     cudaStream_t stream);
int attn_d64_long(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                  int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, void* o_chunks, float* stats,
                  cudaStream_t stream);

// xattn.cu -----------------------------------------------------------------------------------------
int xattn_fused(const void* xn, const void* kp, const void* vt, const float* bias, float* h, int frames,
                int rows_per_frame, int frame0, int total_frames, cudaStream_t stream);
int xattn_split(const void* xn, const void* kp, const void* vt, const float* bias, float* h, void* probs_f16, int frames,
                int rows_per_frame, int frame0, int total_frames, cudaStream_t stream);
int xattn_fold(const void* ctxkv, const void* wq_t, const void* w_o, int depth, int frames, void* kp, void* vt,
               cudaStream_t stream);

// norm.cu ------------------------------------------------------------------------------------------
int ln_rows(const float* x, int64_t ldx, const float* gamma, const float* beta, int64_t mod_frame_stride,
            int rows_per_frame, int gamma_plus_one, void* out, int64_t ldo, int out_f32, int64_t rows, int D,
            float eps, cudaStream_t stream);

// dit_misc.cu --------------------------------------------------------------------------------------
int dit_mod_table(const float* sigma, int S, const float* freqs, int half, const float* map0_w, const float* map0_b,
                  const float* map1_w, const float* map1_b, const float* ada_w, const float* ada_b, int depth,
                  int dim, float* t_emb_ws, float* mod, cudaStream_t stream);
int dit_boundary(const float* h, const float* ln_w, const float* ln_b, const float* w_out_t, const float* w_in_t,
                 const float* x_in, const float* x_base, float* d_buf, float* x_out, float* h_next,
                 const float* sigma, int64_t sigma_stride, const float* sigma_other, int64_t sigma_other_stride,
                 int mode, int rows_per_frame, int C, int64_t T, int dim, float sigma_data, cudaStream_t stream,
                 const void* pack = nullptr);   // pack: dit_boundary_pack's image of the weights (null: packed on the fly)
int dit_boundary_pack(const float* ln_w, const float* ln_b, const float* w_out_t, const float* w_in_t, int C,
                      void* pack, cudaStream_t stream);
int64_t dit_boundary_pack_bytes();
int radar_tokens(const float* feat, int B, int nr, int na, int ne, int cz, const float* w, const float* b,
                 const float* r_emb, const float* a_emb, const float* e_emb, int dim, float* tok_f32, void* tok_bf16,
                 cudaStream_t stream);

// ae.cu / ae_query.cu ---------------------------------------------------------------------------------
int linear_smallk(const float* x, int K, const float* wt, const float* b, float* out, int64_t T, int N,
                  cudaStream_t stream);
int ln_dot_rows(const float* x, const float* g, const float* b, const float* w, float* out, int64_t rows, int D,
                float eps, cudaStream_t stream);
int ae_query(const float* queries, int B, int64_t Q, const void* wpe_bf16, const float* pe_bias, const float* ln_g,
             const float* ln_b, const void* kprime_bf16, const float* vprime, const float* c0, const float* freq24,
             float* logits, int dim, int n_latents, cudaStream_t stream);

// conv3d.cu / enc_misc.cu ------------------------------------------------------------------------------
int conv3d_cl(const void* x_bf16, const void* w_packed, int w_rows, const float* bias, const float* resid, float* out,
              int B, int D, int H, int W, int Cin, int Cout, int stride, cudaStream_t stream);
int enc_conv_in(const float* x, const float* w, const float* bias, float* out, int B, int D, int H, int W, int Cin,
                int Cout, cudaStream_t stream);
int gn_stats(const float* x, int B, int64_t V, int C, int groups, double* stats, cudaStream_t stream);
int gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, void* out_bf16, int B,
             int64_t V, int C, int groups, float eps, int mode, cudaStream_t stream);
int enc_attn(const float* qkv, void* out_bf16, int B, int n, int C, cudaStream_t stream);

}  // namespace rald
