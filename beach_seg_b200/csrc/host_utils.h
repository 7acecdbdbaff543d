// Host-side plumbing shared by the C-ABI translation units: error reporting, TMA descriptor
// encoding through the driver entry point (no link-time libcuda dependency), launch checks.
#pragma once
#include <utility>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace bseg {

// thread-local last-error string behind bseg_last_error()
char* last_error_buf();
void set_error(const char* fmt, ...);

#define BSEG_CHECK_CUDA(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::bseg::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -static_cast<int>(_e);                                                             \
    }                                                                                           \
  } while (0)

#define BSEG_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      ::bseg::set_error(__VA_ARGS__);    \
      return -1000;                      \
    }                                    \
  } while (0)

// Encode a tiled TMA descriptor for a bf16 tensor with `rank` dims (dim 0 innermost, contiguous).
// strides_bytes[i] is the byte stride of dim i+1 (rank-1 entries). 128B swizzle, zero OOB fill.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);

// The same for fp32 tensors, no swizzle (dense rows of box[0] floats in shared memory).
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

inline int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t ld_elems,
                             uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {ld_elems * 2};
  uint32_t box[2] = {box_inner, box_outer};
  return make_tmap_bf16(out, base, 2, dims, strides, box);
}

int num_sms();  // of the current device

// Function attributes (opt-in dynamic shared memory), cluster occupancy and SM counts are PER DEVICE, and the Python API
// accepts any device=: call-site caches are therefore keyed by the current device ordinal.
constexpr int kMaxDevices = 64;
int current_device();
struct PerDeviceFlag {
  bool done[kMaxDevices] = {};
  bool first() {  // true the first time this call site runs on the current device
    const int d = current_device();
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
struct PerDeviceInt {
  int v[kMaxDevices] = {};
  int& get() { return v[current_device()]; }
};

// Launch with (pdl_set(1), the default) or without the programmatic-stream-serialization attribute; only for kernels
// that call pdl_wait() (common.cuh).  Hides the launch latency and the prologue of a kernel under the tail of its
// predecessor: what a batch-1 forward (181 launches of 10-40 us) is made of.
int pdl_set(int on);  // < 0 queries; returns the previous setting
// The attribute is only attached inside a PdlScope(true) of the calling thread: bseg_forward* opens one for launches of
// at most kPdlMaxBatch tiles (measured on B200: batch 1 -11 %, batch 2 -7 %, batch 4 and 16 unchanged, batch 64 +2 %:
// early-scheduled dependents take SM slots from a predecessor that still has work for them).
constexpr int kPdlMaxBatch = 2;
bool pdl_active();
struct PdlScope {
  explicit PdlScope(bool on);
  ~PdlScope();
  bool prev_;
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_active() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// launch accounting behind bseg_launch_count()
void count_launch(int n = 1);
long long launch_count();

// Optional per-launch CUDA-event timing on the launching stream (bseg_profile_*), grouped by kernel category.
enum ProfCat : int {
  CAT_GEMM = 0, CAT_ATTENTION, CAT_LAYERNORM, CAT_DECODER_HEAD, CAT_INGEST, CAT_DECODE, CAT_VOTE, CAT_ELEMENTWISE,
  CAT_LOSS, CAT_COUNT
};
bool prof_enabled();
void prof_set_enabled(bool on);
void prof_collect(double* ms, long long* launches, double* work, double* bytes);  // arrays of CAT_COUNT
void prof_collect_sub(double* ms, double* work);  // arrays of 16: the last collect()'s per-`sub` split (GEMM epilogue modes)
struct ProfScope {
  ProfScope(int cat, double work, double bytes, cudaStream_t stream, int sub = 0);
  ~ProfScope();
  int idx_;
  cudaStream_t stream_;
};

}  // namespace bseg
