#include "gemm.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

template <int BLOCK_N, int MODE>
static int launch_gemm_t(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long M, int N, int K,
                         const GemmEpiParams& ep, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(lda),
                             GEMM_BLOCK_K, GEMM_BLOCK_M);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, W, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                         GEMM_BLOCK_K, BLOCK_N);
  if (rc) return rc;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, MODE>;
  static bool attr_set = false;  // per template instantiation
  if (!attr_set) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const long long tiles = ((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M) * (N / BLOCK_N);
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  ProfScope prof(CAT_GEMM, 2.0 * static_cast<double>(M) * N * K,
                 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K + static_cast<double>(M) * N), stream,
                 MODE + (K > 2048 ? 8 : 0));
  kern<<<grid, GEMM_THREADS, Cfg::kSmemBytes, stream>>>(ta, tb, M, N, K, ep);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int launch_gemm(int mode, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long M, int N, int K,
                const GemmEpiParams& ep, cudaStream_t stream) {
  BSEG_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%lld N=%d K=%d", M, N, K);
  BSEG_REQUIRE(K % GEMM_BLOCK_K == 0, "gemm: K=%d must be a multiple of %d", K, GEMM_BLOCK_K);
  BSEG_REQUIRE(N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
  BSEG_REQUIRE((lda * 2) % 16 == 0, "gemm: lda=%lld violates TMA 16-byte stride alignment", lda);
  BSEG_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  const bool wide = (N % 256 == 0);
#define BSEG_GEMM_CASE(MODE_)                                                               \
  case MODE_:                                                                               \
    return wide ? launch_gemm_t<256, MODE_>(A, lda, W, M, N, K, ep, stream)                 \
                : launch_gemm_t<128, MODE_>(A, lda, W, M, N, K, ep, stream);
  switch (mode) {
    BSEG_GEMM_CASE(EPI_BF16)
    BSEG_GEMM_CASE(EPI_BF16_GELU)
    BSEG_GEMM_CASE(EPI_F32)
    BSEG_GEMM_CASE(EPI_RESID_F32)
    BSEG_GEMM_CASE(EPI_QKV)
    BSEG_GEMM_CASE(EPI_EMBED)
    BSEG_GEMM_CASE(EPI_PIXSHUF)
    default:
      BSEG_REQUIRE(false, "gemm: unknown epilogue mode %d", mode);
  }
#undef BSEG_GEMM_CASE
  return 0;
}

}  // namespace bseg
