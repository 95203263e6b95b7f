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

}  // namespace rald
