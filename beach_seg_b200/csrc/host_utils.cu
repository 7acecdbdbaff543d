#include "host_utils.h"

#include <atomic>
#include <mutex>
#include <vector>

namespace bseg {

char* last_error_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 1024, fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  BSEG_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base),
                  gdims, gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BSEG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed: CUresult=%d (rank=%d dims0=%llu box0=%u)",
               static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]), box[0]);
  return 0;
}

int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  EncodeTiledFn fn = get_encode_fn();
  BSEG_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdims,
                  gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BSEG_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(f32) failed: CUresult=%d (rank=%d dims0=%llu box0=%u)",
               static_cast<int>(r), rank, static_cast<unsigned long long>(dims[0]), box[0]);
  return 0;
}

static std::atomic<int> g_pdl{1};
int pdl_set(int on) {
  const int prev = g_pdl.load(std::memory_order_relaxed);
  if (on >= 0) g_pdl.store(on != 0, std::memory_order_relaxed);
  return prev;
}
static thread_local bool t_pdl_scope = false;
bool pdl_active() { return t_pdl_scope && g_pdl.load(std::memory_order_relaxed) != 0; }
PdlScope::PdlScope(bool on) : prev_(t_pdl_scope) { t_pdl_scope = on; }
PdlScope::~PdlScope() { t_pdl_scope = prev_; }
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

namespace {
struct ProfRec { int cat; cudaEvent_t a, b; double work, bytes; int sub; };
double g_sub_ms[16], g_sub_work[16];
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::mutex g_prof_mu;
}  // namespace
bool prof_enabled() { return g_prof_on; }
void prof_set_enabled(bool on) { g_prof_on = on; }
ProfScope::ProfScope(int cat, double work, double bytes, cudaStream_t stream, int sub) : idx_(-1), stream_(stream) {
  if (!g_prof_on) return;
  ProfRec r{cat, nullptr, nullptr, work, bytes, sub & 15};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, stream);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  idx_ = static_cast<int>(g_prof.size());
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  cudaEventRecord(g_prof[idx_].b, stream_);
}
void prof_collect(double* ms, long long* launches, double* work, double* bytes) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (int i = 0; i < CAT_COUNT; ++i) { ms[i] = 0; launches[i] = 0; work[i] = 0; bytes[i] = 0; }
  for (int i = 0; i < 16; ++i) { g_sub_ms[i] = 0; g_sub_work[i] = 0; }
  for (auto& r : g_prof) {
    cudaEventSynchronize(r.b);
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms[r.cat] += t;
    if (r.cat == CAT_GEMM) { g_sub_ms[r.sub] += t; g_sub_work[r.sub] += r.work; }
    launches[r.cat] += 1;
    work[r.cat] += r.work;
    bytes[r.cat] += r.bytes;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_prof.clear();
}

void prof_collect_sub(double* ms, double* work) {
  for (int i = 0; i < 16; ++i) { ms[i] = g_sub_ms[i]; work[i] = g_sub_work[i]; }
}

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

int num_sms() {
  static PerDeviceInt cache;
  int& n = cache.get();
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device());
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace bseg
