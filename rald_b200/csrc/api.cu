// C ABI of librald_b200.so (declared in include/rald_b200.h). Plain pointers and sizes only.
#include "../../include/rald_b200.h"

#include "host.cuh"
#include "kernels.h"

extern "C" {

int rald_abi_version(void) { return RALD_ABI_VERSION; }

const char* rald_last_error(void) { return rald::get_error(); }

int rald_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                   const float* bias, const float* resid, int64_t ldr, int M, int N, int K, int out_mode,
                   int bn_hint, void* stream) {
  return rald::gemm_bf16(A, lda, W, ldw, out, ldo, bias, resid, ldr, M, N, K, out_mode, bn_hint,
                         static_cast<cudaStream_t>(stream));
}

int rald_gemm_bf16_accum(const void* A, int64_t lda, const void* W, int64_t ldw, float* out, int64_t ldo, int M, int N,
                         int K, void* stream) {
  return rald::gemm_bf16_accum_splitk(A, lda, W, ldw, 0, out, ldo, M, N, K, static_cast<cudaStream_t>(stream));
}

int rald_gemm_bf16_accum_shift(const void* A, int64_t lda, const void* W, int64_t ldw, int w_col_shift, float* out,
                               int64_t ldo, int M, int N, int K, void* stream) {
  return rald::gemm_bf16_accum_splitk(A, lda, W, ldw, w_col_shift, out, ldo, M, N, K, static_cast<cudaStream_t>(stream));
}

int rald_gemm_bf16_accum_taps(const void* A, int64_t lda, const void* W, int64_t ldw, int w_rows, int n_taps,
                              const int* tap_shifts_host, int panel_len, int w_halo, float* out, int64_t ldo, int M, int K,
                              void* stream) {
  return rald::gemm_bf16_accum_taps(A, lda, W, ldw, w_rows, n_taps, tap_shifts_host, panel_len, w_halo, out, ldo, M, K,
                                    static_cast<cudaStream_t>(stream));
}

int rald_gemm_bf16_f16cols(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                           const float* bias, int M, int N, int K, int f16_start, int f16_period, void* stream) {
  return rald::gemm_bf16_f16cols(A, lda, W, ldw, out, ldo, bias, M, N, K, f16_start, f16_period,
                                 static_cast<cudaStream_t>(stream));
}

int rald_gemm_bf16_wsplit(const void* A, int64_t lda, const void* W_hilo, int64_t ldw, void* out, int64_t ldo,
                          const float* bias, const float* resid, int64_t ldr, int M, int N, int K, int out_mode,
                          int f16_start, int f16_period, int gelu_exact, void* stream) {
  return rald::gemm_bf16_wsplit(A, lda, W_hilo, ldw, out, ldo, bias, resid, ldr, M, N, K, out_mode, f16_start,
                                f16_period, gelu_exact, static_cast<cudaStream_t>(stream));
}

int rald_attn_d64(const void* Q, int64_t ldq, const void* K, int64_t ldk, const void* V, int64_t ldv, void* O,
                  int64_t ldo, int frames, int heads, int Sq, int Skv, float scale, void* stream) {
  return rald::attn_d64(Q, ldq, K, ldk, V, ldv, O, ldo, frames, heads, Sq, Skv, scale,
                        static_cast<cudaStream_t>(stream));
}

int rald_ln_rows(const float* x, int64_t ldx, const float* gamma, const float* beta, int64_t mod_frame_stride,
                 int rows_per_frame, int gamma_plus_one, void* out, int64_t ldo, int out_f32, int64_t rows, int D,
                 float eps, void* stream) {
  return rald::ln_rows(x, ldx, gamma, beta, mod_frame_stride, rows_per_frame, gamma_plus_one, out, ldo, out_f32, rows,
                       D, eps, static_cast<cudaStream_t>(stream));
}

int rald_dit_mod_table(const float* sigma, int S, const float* freqs, int half, const float* map0_w,
                       const float* map0_b, const float* map1_w, const float* map1_b, const float* ada_w,
                       const float* ada_b, int depth, int dim, float* t_emb_ws, float* mod, void* stream) {
  return rald::dit_mod_table(sigma, S, freqs, half, map0_w, map0_b, map1_w, map1_b, ada_w, ada_b, depth, dim,
                             t_emb_ws, mod, static_cast<cudaStream_t>(stream));
}

int rald_dit_boundary(const float* h, const float* ln_w, const float* ln_b, const float* w_out_t,
                      const float* w_in_t, const float* x_in, const float* x_base, float* d_buf, float* x_out,
                      float* h_next, const float* sigma, int64_t sigma_stride, const float* sigma_other,
                      int64_t sigma_other_stride, int mode, int rows_per_frame, int C, int64_t T, int dim,
                      float sigma_data, const void* pack, void* stream) {
  return rald::dit_boundary(h, ln_w, ln_b, w_out_t, w_in_t, x_in, x_base, d_buf, x_out, h_next, sigma, sigma_stride,
                            sigma_other, sigma_other_stride, mode, rows_per_frame, C, T, dim, sigma_data,
                            static_cast<cudaStream_t>(stream), pack);
}

int64_t rald_dit_boundary_pack_bytes(void) { return rald::dit_boundary_pack_bytes(); }

int rald_dit_boundary_pack(const float* ln_w, const float* ln_b, const float* w_out_t, const float* w_in_t, int C,
                           void* pack, void* stream) {
  return rald::dit_boundary_pack(ln_w, ln_b, w_out_t, w_in_t, C, pack, static_cast<cudaStream_t>(stream));
}

int rald_radar_tokens(const float* feat, int B, int nr, int na, int ne, int cz, const float* w, const float* b,
                      const float* r_emb, const float* a_emb, const float* e_emb, int dim, float* tok_f32,
                      void* tok_bf16, void* stream) {
  return rald::radar_tokens(feat, B, nr, na, ne, cz, w, b, r_emb, a_emb, e_emb, dim, tok_f32, tok_bf16,
                            static_cast<cudaStream_t>(stream));
}

}  // extern "C"
