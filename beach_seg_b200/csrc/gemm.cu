#include "gemm.cuh"
#include "host_utils.h"
#include "kernels.h"

#include <stdlib.h>

namespace bseg {

// BSEG_GEMM_2CTA=0/1 selects the CTA-pair kernel (cta_group::2, 256-row tiles) for the N % 256 == 0 GEMMs.
static int g_cta_pairs = -1;  // -1: not decided yet (environment, else the default)
static bool use_cta_pairs() {
  if (g_cta_pairs < 0) {
    const char* e = getenv("BSEG_GEMM_2CTA");
    g_cta_pairs = e ? (atoi(e) != 0) : 1;  // default: CTA pairs (7-13 % faster on the layer GEMMs, bit-identical)
  }
  return g_cta_pairs != 0;
}
static int g_small_tiles = 1;  // 128 x 128 tiles for launches too small to fill the SMs with 256-wide ones
int gemm_set_small_tiles(int on) {
  const int prev = g_small_tiles;
  if (on >= 0) g_small_tiles = on != 0;
  return prev;
}
static int g_fused_ln = 0;  // 0: off (default), 1: lin2 also emits the next norm1 (EPI_RESID_LN), 2: proj emits norm2 as well
int gemm_set_fused_ln(int level) {
  const int prev = g_fused_ln;
  if (level >= 0) g_fused_ln = level > 2 ? 2 : level;
  return prev;
}
int gemm_set_cta_pairs(int on) {
  const int prev = use_cta_pairs() ? 1 : 0;
  if (on >= 0) g_cta_pairs = on != 0;
  return prev;
}

// The CTA-pair variant: clusters of two CTAs, each pair owns 256 x BLOCK_N output tiles.
template <int BLOCK_N, int MODE>
static int launch_gemm_pair_t(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, const GemmRows& gr, int N,
                              int K, const GemmEpiParams& ep, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, 2>;
  CUtensorMap ta, tb;
  int rc;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(gr.rows_per_batch),
                        static_cast<uint64_t>(gr.nbatch)};
    uint64_t strides[2] = {static_cast<uint64_t>(lda) * 2, static_cast<uint64_t>(gr.rows_per_batch) * lda * 2};
    uint32_t box[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
    rc = make_tmap_bf16(&ta, A, 3, dims, strides, box);
  }
  if (rc) return rc;
  const long long M = static_cast<long long>(gr.nbatch) * gr.rows;
  rc = make_tmap_bf16_2d(&tb, W, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                         GEMM_BLOCK_K, Cfg::kBRows);
  if (rc) return rc;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, MODE, 2>;
  static PerDeviceFlag attr_once;  // per template instantiation
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const long long tiles =
      static_cast<long long>(gr.nbatch) * ((gr.rows + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M)) * (N / BLOCK_N);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // (the kernel calls pdl_wait() after its prologue)
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent kernel: launch exactly as many pairs as can be co-resident (normally one per TPC = SMs / 2)
  static PerDeviceInt max_pairs_cache;  // per template instantiation and device
  int& max_pairs = max_pairs_cache.get();
  if (max_pairs == 0) {
    cfg.gridDim = dim3(static_cast<unsigned>(num_sms() & ~1));
    int n = 0;
    BSEG_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    BSEG_REQUIRE(n > 0, "gemm: no CTA pair of %d B shared memory can be resident", Cfg::kSmemBytes);
    max_pairs = n < num_sms() / 2 ? n : num_sms() / 2;
  }
  const long long pairs = max_pairs;
  cfg.gridDim = dim3(static_cast<unsigned>(2 * (tiles < pairs ? tiles : pairs)));
  ProfScope prof(CAT_GEMM, 2.0 * static_cast<double>(M) * N * K,
                 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K + static_cast<double>(M) * N), stream,
                 MODE == EPI_RESID_LN ? (K > 2048 ? 13 : 12) : MODE + (K > 2048 ? 8 : 0));
  cfg.numAttrs = pdl_active() ? 2 : 1;
  BSEG_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, gr, N, K, ep));
  count_launch();
  return 0;
}

template <int BLOCK_N, int MODE>
static int launch_gemm_t(const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, const GemmRows& gr, int N, int K,
                         const GemmEpiParams& ep, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  CUtensorMap ta, tb;
  int rc;
  {
    uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(gr.rows_per_batch),
                        static_cast<uint64_t>(gr.nbatch)};
    uint64_t strides[2] = {static_cast<uint64_t>(lda) * 2, static_cast<uint64_t>(gr.rows_per_batch) * lda * 2};
    uint32_t box[3] = {GEMM_BLOCK_K, GEMM_BLOCK_M, 1};
    rc = make_tmap_bf16(&ta, A, 3, dims, strides, box);
  }
  if (rc) return rc;
  const long long M = static_cast<long long>(gr.nbatch) * gr.rows;
  rc = make_tmap_bf16_2d(&tb, W, static_cast<uint64_t>(K), static_cast<uint64_t>(N), static_cast<uint64_t>(K),
                         GEMM_BLOCK_K, BLOCK_N);
  if (rc) return rc;
  auto kern = gemm_bf16_tcgen05_kernel<BLOCK_N, MODE>;
  static PerDeviceFlag attr_once;  // per template instantiation
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  const long long tiles = static_cast<long long>(gr.nbatch) * ((gr.rows + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M) * (N / BLOCK_N);
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  ProfScope prof(CAT_GEMM, 2.0 * static_cast<double>(M) * N * K,
                 2.0 * (static_cast<double>(M) * K + static_cast<double>(N) * K + static_cast<double>(M) * N), stream,
                 MODE == EPI_RESID_LN ? (K > 2048 ? 13 : 12) : MODE + (K > 2048 ? 8 : 0));
  BSEG_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(GEMM_THREADS), Cfg::kSmemBytes, stream, ta, tb, gr, N, K, ep));
  count_launch();
  return 0;
}

int launch_gemm(int mode, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long M, int N, int K,
                const GemmEpiParams& ep, cudaStream_t stream) {
  BSEG_REQUIRE(M > 0 && M < (1ll << 31), "gemm: M=%lld out of range", M);
  GemmRows gr{M, 1, 0, static_cast<int>(M)};
  return launch_gemm_rows(mode, A, lda, W, gr, N, K, ep, stream);
}

int launch_gemm_rows(int mode, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, const GemmRows& gr_in,
                     int N, int K, const GemmEpiParams& ep, cudaStream_t stream) {
  GemmRows gr = gr_in;
  // a range that covers every batch entry completely is one contiguous block of rows: no per-entry tile padding
  // (1568 rows = 12.25 tiles of 128) and 256-row pair tiles fit (B * 1568 is a multiple of 256 for even B)
  if (gr.nbatch > 1 && gr.row_begin == 0 && gr.rows == gr.rows_per_batch &&
      gr.rows_per_batch * gr.nbatch < (1ll << 31)) {
    gr = GemmRows{gr.rows_per_batch * gr.nbatch, 1, 0, static_cast<int>(gr.rows_per_batch * gr.nbatch)};
  }
  BSEG_REQUIRE(gr.rows > 0 && gr.nbatch > 0 && N > 0 && K > 0, "gemm: empty problem rows=%d N=%d K=%d", gr.rows, N, K);
  BSEG_REQUIRE(gr.row_begin >= 0 && gr.row_begin + gr.rows <= gr.rows_per_batch, "gemm: row range outside the batch");
  BSEG_REQUIRE(K % GEMM_BLOCK_K == 0, "gemm: K=%d must be a multiple of %d", K, GEMM_BLOCK_K);
  BSEG_REQUIRE(N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
  BSEG_REQUIRE((lda * 2) % 16 == 0, "gemm: lda=%lld violates TMA 16-byte stride alignment", lda);
  BSEG_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: operands must be 16-byte aligned");
  bool wide = (N % 256 == 0);
  // 256-row tiles pad a ragged row range more than 128-row tiles do (the decoder's 812-row query-half slices: 1024
  // instead of 896 rows computed); pairs only when that costs less than they gain
  const int t128 = (gr.rows + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M, t256 = (gr.rows + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
  bool pairs = use_cta_pairs() && 2 * t256 * 100 <= t128 * 105;
  // Small launches (one or a few tiles of a scene: M = 1568 .. 6272 rows) leave most SMs without a tile when the tiles
  // are 256 wide: 7 x 4 pair tiles for the N = 1024 GEMMs of a single tile.  Estimate each variant as
  // (waves of tiles over its slots) x (tensor cycles per K step of its tile) and take 128 x 128 one-CTA tiles when
  // they win clearly; all variants accumulate over K in the same order, so the results are bit-identical.
  if (wide && g_small_tiles != 0) {
    const long long sms = num_sms();
    const long long nb = gr.nbatch;
    const long long tiles_pair = nb * t256 * (N / 256), tiles_256 = nb * t128 * (N / 256), tiles_128 = nb * t128 * (N / 128);
    const long long slots_pair = sms / 2 > 0 ? sms / 2 : 1;
    const long long est_wide = pairs ? ((tiles_pair + slots_pair - 1) / slots_pair) * 128
                                     : ((tiles_256 + sms - 1) / sms) * 128;
    const long long est_128 = ((tiles_128 + sms - 1) / sms) * 90;  // measured: a 128 x 128 tile costs ~0.7 of a 256-wide one
    if (est_128 * 100 <= est_wide * 85) wide = false;
  }
  if (mode == EPI_RESID_LN) {
    // residual + LayerNorm: the four N tiles of a row block exchange row statistics, eight 128-column slices per row
    BSEG_REQUIRE(N == 1024 && gr.nbatch == 1 && gr.row_begin == 0,
                 "gemm: the residual+LayerNorm epilogue needs N == 1024 and one contiguous row range");
    BSEG_REQUIRE(ep.ln_gamma && ep.ln_beta && ep.ln_out && ep.ln_stats && ep.ln_tag < 255 && ep.resid && ep.out,
                 "gemm: residual+LayerNorm epilogue with a null pointer");
    BSEG_REQUIRE(ep.ld_ln == 1024 && ep.ldc == 1024 && ep.ldr == 1024, "gemm: residual+LayerNorm operands must be dense [M,1024]");
    if (pairs) return launch_gemm_pair_t<256, EPI_RESID_LN>(A, lda, W, gr, N, K, ep, stream);
    return launch_gemm_t<256, EPI_RESID_LN>(A, lda, W, gr, N, K, ep, stream);
  }
#define BSEG_GEMM_CASE(MODE_)                                                               \
  case MODE_:                                                                               \
    if (wide && pairs) return launch_gemm_pair_t<256, MODE_>(A, lda, W, gr, N, K, ep, stream); \
    return wide ? launch_gemm_t<256, MODE_>(A, lda, W, gr, N, K, ep, stream)                 \
                : launch_gemm_t<128, MODE_>(A, lda, W, gr, N, K, ep, stream);
  switch (mode) {
    BSEG_GEMM_CASE(EPI_BF16)
    BSEG_GEMM_CASE(EPI_BF16_GELU)
    BSEG_GEMM_CASE(EPI_F32)
    BSEG_GEMM_CASE(EPI_RESID_F32)
    BSEG_GEMM_CASE(EPI_QKV)
    BSEG_GEMM_CASE(EPI_EMBED)
    BSEG_GEMM_CASE(EPI_PIXSHUF)
    BSEG_GEMM_CASE(EPI_DGELU)
    default:
      BSEG_REQUIRE(false, "gemm: unknown epilogue mode %d", mode);
  }
#undef BSEG_GEMM_CASE
  return 0;
}

}  // namespace bseg
