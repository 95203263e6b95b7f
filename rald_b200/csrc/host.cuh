// Host-side plumbing shared by all translation units: error reporting for the C ABI and
// TMA tensor-map encoding (cuTensorMapEncodeTiled resolved through the runtime so the library
// links without libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace rald {

// thread-local last error string, read by rald_last_error()
void set_error(const char* fmt, ...);
const char* get_error();

#define RALD_CHECK_CUDA(expr)                                                                        \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ::rald::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return -static_cast<int>(_e) - 1000;                                                           \
    }                                                                                                \
  } while (0)

#define RALD_REQUIRE(cond, ...)                  \
  do {                                           \
    if (!(cond)) {                               \
      ::rald::set_error(__VA_ARGS__);            \
      return -1;                                 \
    }                                            \
  } while (0)

#define RALD_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != 0) return _r;     \
  } while (0)

// 2-D row-major matrix [rows, cols] of 16-bit elements (cols contiguous, row pitch ld elements),
// box = box_rows x 64 columns, 128-byte swizzle, out-of-bounds reads return zero.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows);
// same for 32-bit elements (tf32 operands): box = box_rows x 32 columns (128 bytes)
int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows);
// Output tiles for TMA stores: box = box_rows (32 or 128) rows x 128 bytes (32 fp32 or 64 bf16 columns), 128-byte swizzle.
int make_tmap_out(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, bool f32,
                  uint32_t box_rows = 32);
// General tiled map over 16-bit elements: dims/strides innermost first (strides in BYTES for dims 1..rank-1).
int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

int device_sm_count();
bool pdl_enabled();  // RALD_B200_PDL != 0 (default on)

// Launch with the programmatic-dependent-launch attribute (the kernel must call pdl_wait() before touching memory
// written by its predecessor in the stream).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                               unsigned cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  return launch_pdl_cluster(kern, grid, block, smem, stream, 1u, static_cast<Args&&>(args)...);
}

// ---- launch accounting (rald_launch_count) and optional per-launch CUDA-event timing (rald_prof_*) ----
enum ProfFamily : int { FAM_GEMM = 0, FAM_ATTN = 1, FAM_LN = 2, FAM_BOUNDARY = 3, FAM_CONV3D = 4, FAM_GN = 5,
                        FAM_AE_QUERY = 6, FAM_OTHER = 7, FAM_FPS = 8, FAM_XATTN = 9, FAM_COUNT = 10 };
void count_launch();
// While alive, brackets the launches issued on `stream` with a pair of CUDA events if timing of `family` is
// enabled (bench.py's roofline leg); `work` = algorithmic flops or bytes of the bracketed launch.
struct ProfScope {
  ProfScope(int family, cudaStream_t stream, double work);
  ~ProfScope();
  int slot;
  cudaStream_t stream;
};

#define RALD_LAUNCHED()                     \
  do {                                      \
    RALD_CHECK_CUDA(cudaGetLastError());    \
    ::rald::count_launch();                 \
  } while (0)

}  // namespace rald
