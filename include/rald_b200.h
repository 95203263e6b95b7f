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

#define RALD_ABI_VERSION 1

int rald_abi_version(void);
const char* rald_last_error(void);

/* out = epilogue(A[M,K] @ W[N,K]^T), A and W bf16, fp32 accumulation on tcgen05 tensor cores.
 *   out_mode 0: out bf16 [M,N] (+bias)
 *   out_mode 1: out f32  [M,N] = acc (+bias) (+resid f32 [M,ldr]); out may alias resid
 *   out_mode 2: GEGLU, out bf16 [M,N/2]; W rows / bias packed in groups of 32 = 16 value + 16 gate rows
 * Replaces nn.Linear (+ residual add / GEGLU) at model/models_radar_generation.py:58-64, 76, 91-95, 113,
 * 166-168 and model/models_ae.py:60-62, 87-89, 105. bn_hint: 0 = auto, else 32/64/128/256 (tile N). */
int rald_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* out, int64_t ldo,
                   const float* bias, const float* resid, int64_t ldr, int M, int N, int K, int out_mode,
                   int bn_hint, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RALD_B200_H */
