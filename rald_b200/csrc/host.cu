#include "host.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "../../include/rald_b200.h"

namespace rald {

static thread_local char g_err[1024] = {0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Descriptor cache: a tensor map is a pure function of (data type, base pointer, dims, strides, box,
// element strides) — no device state — so the encoded 128 bytes are memoised under exactly that key; a buffer that is
// freed and reallocated at the same address with the same geometry gets the same (still valid) descriptor. The hot
// loops re-present the same few hundred (pointer, shape) pairs on every call (workspaces and packed weights), so after
// the first evaluation no launch pays for cuTensorMapEncodeTiled (~1.5 us each, 3 per GEMM) any more.
struct TmapKey {
  uint64_t w[16];
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t v : k.w) { h ^= v; h *= 0x100000001b3ull; h ^= h >> 29; }
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t TMAP_CACHE_MAX = 8192;
static std::atomic<uint64_t> g_tmap_hits{0}, g_tmap_misses{0};

static int encode_uncached(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);

static int encode(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  RALD_REQUIRE(rank >= 1 && rank <= 5, "TMA rank %d unsupported", rank);
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = reinterpret_cast<uint64_t>(base);
  key.w[1] = ((uint64_t)dt << 8) | (uint64_t)rank;
  for (int i = 0; i < rank; ++i) {
    key.w[2 + i] = dims[i];
    key.w[11 + i] = ((uint64_t)box[i] << 32) | (uint64_t)(elem_strides ? elem_strides[i] : 1u);
  }
  for (int i = 0; i + 1 < rank; ++i) key.w[7 + i] = strides_bytes[i];
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
      return 0;
    }
  }
  RALD_TRY(encode_uncached(out, dt, base, rank, dims, strides_bytes, box, elem_strides));
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmap_cache.size() >= TMAP_CACHE_MAX) g_tmap_cache.clear();
  g_tmap_cache.emplace(key, *out);
  return 0;
}

static int encode_uncached(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                           const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = get_encode_fn();
  RALD_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  RALD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    RALD_REQUIRE((gstr[i] & 15) == 0, "TMA stride %llu (dim %d) not a multiple of 16 bytes",
                 (unsigned long long)gstr[i], i + 1);
  }
  CUresult r = fn(out, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RALD_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu x %llu)",
               (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0));
  return 0;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * 2};
  uint32_t box[2] = {64, box_rows};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, 2, dims, strides, box, nullptr);
}

int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                     uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * 4};
  uint32_t box[2] = {32, box_rows};
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, 2, dims, strides, box, nullptr);
}

int make_tmap_out(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, bool f32,
                  uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t strides[1] = {ld * (f32 ? 4u : 2u)};
  uint32_t box[2] = {f32 ? 32u : 64u, box_rows};
  return encode(out, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, 2, dims, strides,
                box, nullptr);
}

int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, elem_strides);
}

// ---- launch accounting / per-launch timing -------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { cudaEvent_t e0, e1; int family; double work; };
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;      // records of the current session
static std::vector<cudaEvent_t> g_pool;  // recycled events
static unsigned g_prof_mask = 0;
constexpr size_t PROF_MAX = 1 << 16;

static cudaEvent_t pool_get() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(int family, cudaStream_t st, double work) : slot(-1), stream(st) {
  if (!(g_prof_mask & (1u << family))) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof.size() >= PROF_MAX) return;
  ProfRec r{pool_get(), pool_get(), family, work};
  if (r.e0 == nullptr || r.e1 == nullptr) return;
  cudaEventRecord(r.e0, st);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[slot].e1, stream);
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("RALD_B200_PDL");
    on = (e == nullptr || e[0] != '0') ? 1 : 0;
  }
  return on == 1;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

}  // namespace rald

extern "C" uint64_t rald_launch_count(void) { return rald::g_launches.load(); }
extern "C" int rald_tmap_cache_stats(uint64_t* hits, uint64_t* misses) {
  if (hits) *hits = rald::g_tmap_hits.load();
  if (misses) *misses = rald::g_tmap_misses.load();
  return 0;
}
extern "C" void rald_launch_count_add(uint64_t n) { rald::g_launches.fetch_add(n); }

extern "C" int rald_prof_enable(unsigned family_mask) {
  std::lock_guard<std::mutex> lk(rald::g_prof_mu);
  for (auto& r : rald::g_prof) { rald::g_pool.push_back(r.e0); rald::g_pool.push_back(r.e1); }
  rald::g_prof.clear();
  rald::g_prof_mask = family_mask;
  return 0;
}

extern "C" int rald_prof_collect(int family, double* total_ms, double* total_work, int64_t* launches) {
  using namespace rald;
  RALD_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  double ms = 0.0, work = 0.0;
  int64_t n = 0;
  for (auto& r : g_prof) {
    if (r.family != family) continue;
    float t = 0.f;
    RALD_CHECK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t; work += r.work; ++n;
  }
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = work;
  if (launches) *launches = n;
  return 0;
}

// Per-record dump of one family: ms[i], work[i] for up to `cap` launches (returns the number written).
extern "C" int64_t rald_prof_dump(int family, float* ms, double* work, int64_t cap) {
  using namespace rald;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int64_t n = 0;
  for (auto& r : g_prof) {
    if (r.family != family || n >= cap) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) return -1;
    ms[n] = t; work[n] = r.work; ++n;
  }
  return n;
}
