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

}  // extern "C"
