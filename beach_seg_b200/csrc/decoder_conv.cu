// SegGptDecoderHead (modeling_seggpt.py:531-552): conv3x3(64->64, pad 1) -> LayerNorm over C (channels_first,
// eps 1e-6) -> erf-GELU -> conv1x1(64->3), fused into one persistent implicit-GEMM kernel.
//   input  : NHWC bf16 [B, H, W, 64]  (written by the decoder_embed GEMM's pixel-shuffle epilogue)
//   output : NCHW fp32 [B, 3, H, W]   (== pred_masks)
// A pixel tile is 2 image rows x 64 columns = 128 GEMM rows; each of the 9 taps is one K=64 block.  Per column shift
// dx one TMA box brings the 4 input rows y-1 .. y+2 (4 x 8 KB row segments, out-of-bounds rows/columns zero filled ==
// the conv's zero padding); the A tile of tap (dy, dx) is the 16 KB window starting at segment dy, so a tile costs
// 3 x 32 KB of L2->smem traffic instead of 9 x 16 KB (the kernel was bound by that traffic, not by MMA or epilogue).
// N = 64 output channels fit one accumulator tile, so LN + GELU + the 1x1 head run in the epilogue registers.
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

namespace dconv {
// epilogue variants of the same implicit-GEMM main loop
enum : int {
  kModeHead = 0,     // forward: + conv bias -> LN(C) -> GELU -> 1x1 head -> pred fp32 NCHW
  kModeHeadBwd = 1,  // backward through head / GELU / LN (conv output recomputed): d(conv out) -> bf16 NHWC
  kModeDgrad = 2,    // conv3x3 dgrad (flipped, transposed taps): d(conv in) -> bf16 rows of the decoder_embed dgrad operand
};
struct Params {
  const float* conv_b;
  const float* ln_w;
  const float* ln_b;
  const float* head_w;
  const float* head_b;
  float* pred;             // kModeHead: out [B,3,H,W]
  const float* d_pred;     // kModeHeadBwd: in  [B,3,H,W]
  __nv_bfloat16* out_bf16; // kModeHeadBwd: [B, H - y_out0, W, 64];  kModeDgrad: [B*T, 16384] (pixel-unshuffled)
  int B, H, W;
  int ty_begin;            // first tile row (in units of kTileH image rows) this launch computes
  int y_in0;               // image row that coordinate 0 of the input tensor map corresponds to (rows above read as 0)
  int y_out0;              // kModeHeadBwd: image row stored at row 0 of out_bf16
  float eps;
};
constexpr int kTileW = 64, kTileH = 2;
constexpr int kStages = 3;
constexpr int kSegBytes = kTileW * 128;  // 8192: one image row segment of the tile (64 pixels x 64 channels bf16)
constexpr int kABytes = 4 * kSegBytes;   // 32768 per column shift dx: the 4 input rows y-1 .. y+2; the A tile of tap
                                         // (dy, dx) is the 16 KB window starting at segment dy
constexpr int kWBytes = 9 * 64 * 128;    // 73728: 9 taps x [64 out x 64 in] bf16
constexpr int kThreads = 64 + 256;      // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kOffW = 0;
constexpr int kOffA = kOffW + kWBytes;
constexpr int kOffBar = kOffA + kStages * kABytes;
constexpr int kOffPar = kOffBar + 256;   // fp32 params: conv_b[64] ln_w[64] ln_b[64] head_w[192] head_b[3]
constexpr int kOffXch = kOffPar + 400 * 4;  // epilogue exchange between the two channel halves: 2*128*2 + 2*128*4 floats
constexpr int kSmemBytes = kOffXch + (2 * 128 * 2 + 2 * 128 * 4) * 4 + 1024;
constexpr uint32_t kTmemCols = 128;
}  // namespace dconv

template <int MODE>
__global__ void __launch_bounds__(dconv::kThreads, 1)
decoder_conv_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                    const dconv::Params prm) {
  pdl_launch_dependents();
  using namespace dconv;
  const int B = prm.B, H = prm.H, W = prm.W;
  const float eps = prm.eps;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem + kOffW;
  uint8_t* sA = smem + kOffA;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* w_full = bars;
  uint64_t* full_bar = bars + 1;                 // [kStages]
  uint64_t* empty_bar = bars + 1 + kStages;      // [kStages]
  uint64_t* tmem_full = bars + 1 + 2 * kStages;  // [2]
  uint64_t* tmem_empty = bars + 3 + 2 * kStages; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * kStages);
  float* sPar = reinterpret_cast<float*>(smem + kOffPar);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_x = W / kTileW, tiles_y = H / kTileH - prm.ty_begin;
  const long long num_tiles = static_cast<long long>(B) * tiles_y * tiles_x;

  if constexpr (MODE != kModeDgrad) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
      sPar[i] = prm.conv_b[i];
      sPar[64 + i] = prm.ln_w[i];
      sPar[128 + i] = prm.ln_b[i];
    }
    for (int i = threadIdx.x; i < 192; i += blockDim.x) sPar[192 + i] = prm.head_w[i];
    if (threadIdx.x < 3) sPar[384 + threadIdx.x] = prm.head_b[threadIdx.x];
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    mbar_init(w_full, 1);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = uniform_u32(*tmem_slot);
  pdl_wait();  // (the prologue above may run under the tail of the previous kernel)

  if (warp == 0) {
    {
      if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_full, kWBytes);
      for (int t = 0; t < 9; t += 3) tma_load_2d(sW + t * 8192, &tmap_w, w_full, 0, t * 64);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tx = static_cast<int>(tile % tiles_x);
        const int ty = static_cast<int>((tile / tiles_x) % tiles_y) + prm.ty_begin;
        const int b = static_cast<int>(tile / (static_cast<long long>(tiles_x) * tiles_y));
        for (int dx = 0; dx < 3; ++dx) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&full_bar[stage], kABytes);
            tma_load_4d(sA + stage * kABytes, &tmap_x, &full_bar[stage], 0, tx * kTileW + dx - 1,
                        ty * kTileH - 1 - prm.y_in0, b);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t w_addr = smem_u32(sW);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + as * 64;
        for (int dx = 0; dx < 3; ++dx) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t a_addr = smem_u32(sA + stage * kABytes);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_ss(d, umma_desc_sw128_kmajor(a_addr + dy * kSegBytes + k * 32),
                             umma_desc_sw128_kmajor(w_addr + (dy * 3 + dx) * 8192 + k * 32), idesc, (dx | dy | k) != 0);
            umma_commit(&empty_bar[stage]);
            if (dx == 2) umma_commit(&tmem_full[as]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ---- epilogue: 8 warps; a pixel's 64 channels are split between two threads (warps w and w + 4 share a TMEM lane
    // quarter).  LayerNorm statistics and the head / LN-backward sums are combined through shared memory. ----
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;  // channels [32 * half, 32 * half + 32)
    const int r = quarter * 32 + lane;
    const int c0 = half * 32;
    float* xchA = reinterpret_cast<float*>(smem + kOffXch);   // [2 halves][128][2]  (mean, M2) of the half
    float* xchB = xchA + 2 * 128 * 2;                         // [2 halves][128][4]  head partials / LN-backward sums
    int as = 0;
    uint32_t aphase = 0;
    const long long plane = static_cast<long long>(H) * W;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int tx = static_cast<int>(tile % tiles_x);
      const int ty = static_cast<int>((tile / tiles_x) % tiles_y) + prm.ty_begin;
      const int b = static_cast<int>(tile / (static_cast<long long>(tiles_x) * tiles_y));
      const int y = ty * kTileH + (r >> 6), x = tx * kTileW + (r & 63);
      float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
      if constexpr (MODE == kModeHeadBwd) {  // does not depend on the accumulator: issue before the wait
        const float* src = prm.d_pred + static_cast<long long>(b) * 3 * plane + static_cast<long long>(y) * W + x;
        dp0 = src[0];
        dp1 = src[plane];
        dp2 = src[2 * plane];
      }
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * 64 + c0;
      float v[32];
      tmem_ld32(taddr, v);
      tmem_ld_wait();
      // accumulator half is in registers: release the TMEM buffer early
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);

      if constexpr (MODE == kModeDgrad) {
        // rows of the decoder_embed dgrad operand: token (y/16, x/16), column ((y%16)*16 + x%16)*64 + c
        __nv_bfloat16* dst = prm.out_bf16 +
                             (static_cast<long long>(b) * ((H >> 4) * (W >> 4)) + (y >> 4) * (W >> 4) + (x >> 4)) * 16384 +
                             (((y & 15) << 4) + (x & 15)) * 64 + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dst + i) =
              make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]), pack_bf16x2(v[i + 4], v[i + 5]),
                         pack_bf16x2(v[i + 6], v[i + 7]));
      } else {
        // two-pass statistics of this half, then Chan's combination with the other half (exact, no cancellation).
        // All per-channel arithmetic runs on pairs of channels (packed fp32 pipe), parameters come as float2 pairs.
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float2 cb = *reinterpret_cast<const float2*>(sPar + c0 + i);
          add_f32x2(v[i], v[i + 1], cb.x, cb.y);
          add_f32x2(s0, s1, v[i], v[i + 1]);
        }
        const float mean_h = (s0 + s1) * (1.0f / 32.0f);
        float q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float d0 = v[i], d1 = v[i + 1];
          add_f32x2(d0, d1, -mean_h, -mean_h);
          uint64_t q = fma_f32x2_raw(pack_f32x2(d0, d1), pack_f32x2(d0, d1), pack_f32x2(q0, q1));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
        }
        const float m2_h = q0 + q1;
        *reinterpret_cast<float2*>(xchA + (half * 128 + r) * 2) = make_float2(mean_h, m2_h);
        named_bar_sync(1, 256);
        const float2 oth = *reinterpret_cast<const float2*>(xchA + ((half ^ 1) * 128 + r) * 2);
        const float mean = 0.5f * (mean_h + oth.x);
        const float dm = mean_h - oth.x;
        const float var = (m2_h + oth.y + 16.0f * dm * dm) * (1.0f / 64.0f);
        const float rstd = rsqrtf(var + eps);
        if constexpr (MODE == kModeHead) {
          uint64_t p0 = 0, p1 = 0, p2 = 0;  // (even, odd) channel partial sums of the three head outputs
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 ga = *reinterpret_cast<const float2*>(sPar + 64 + c0 + i);
            const float2 be = *reinterpret_cast<const float2*>(sPar + 128 + c0 + i);
            float t0 = v[i], t1 = v[i + 1];
            add_f32x2(t0, t1, -mean, -mean);
            mul_f32x2(t0, t1, t0, t1, rstd, rstd);
            uint64_t y = fma_f32x2_raw(pack_f32x2(t0, t1), pack_f32x2(ga.x, ga.y), pack_f32x2(be.x, be.y));
            float g0, g1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(g0), "=f"(g1) : "l"(y));
            gelu_erf_x2(g0, g1);
            const uint64_t g = pack_f32x2(g0, g1);
            const float2 w0 = *reinterpret_cast<const float2*>(sPar + 192 + c0 + i);
            const float2 w1 = *reinterpret_cast<const float2*>(sPar + 256 + c0 + i);
            const float2 w2 = *reinterpret_cast<const float2*>(sPar + 320 + c0 + i);
            p0 = fma_f32x2_raw(g, pack_f32x2(w0.x, w0.y), p0);
            p1 = fma_f32x2_raw(g, pack_f32x2(w1.x, w1.y), p1);
            p2 = fma_f32x2_raw(g, pack_f32x2(w2.x, w2.y), p2);
          }
          float o0a, o0b, o1a, o1b, o2a, o2b;
          asm("mov.b64 {%0, %1}, %2;" : "=f"(o0a), "=f"(o0b) : "l"(p0));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(o1a), "=f"(o1b) : "l"(p1));
          asm("mov.b64 {%0, %1}, %2;" : "=f"(o2a), "=f"(o2b) : "l"(p2));
          const float o0 = o0a + o0b, o1 = o1a + o1b, o2 = o2a + o2b;
          if (half == 1) *reinterpret_cast<float4*>(xchB + (128 + r) * 4) = make_float4(o0, o1, o2, 0.f);
          named_bar_sync(2, 256);
          if (half == 0) {
            const float4 ob = *reinterpret_cast<const float4*>(xchB + (128 + r) * 4);
            float* dst = prm.pred + static_cast<long long>(b) * 3 * plane + static_cast<long long>(y) * W + x;
            dst[0] = (o0 + ob.x) + sPar[384];
            dst[plane] = (o1 + ob.y) + sPar[385];
            dst[2 * plane] = (o2 + ob.z) + sPar[386];
          }
        } else {
          // pred = head(gelu(LN(v))):  dg = head_w^T dpred;  dyl = dg * gelu'(yl);  LN backward over the 64 channels
          float m1 = 0.f, m2 = 0.f;
          float gyv[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float xh = (v[i] - mean) * rstd;
            const float yl = fmaf(xh, sPar[64 + c0 + i], sPar[128 + c0 + i]);
            const float dg = fmaf(dp0, sPar[192 + c0 + i], fmaf(dp1, sPar[256 + c0 + i], dp2 * sPar[320 + c0 + i]));
            const float gy = dg * gelu_erf_grad(yl) * sPar[64 + c0 + i];
            m1 += gy;
            m2 = fmaf(gy, xh, m2);
            v[i] = xh;
            gyv[i] = gy;
          }
          *reinterpret_cast<float2*>(xchB + (half * 128 + r) * 4) = make_float2(m1, m2);
          named_bar_sync(2, 256);
          const float2 om = *reinterpret_cast<const float2*>(xchB + ((half ^ 1) * 128 + r) * 4);
          m1 = (m1 + om.x) * (1.0f / 64.0f);
          m2 = (m2 + om.y) * (1.0f / 64.0f);
          __nv_bfloat16* dst = prm.out_bf16 +
                               ((static_cast<long long>(b) * (H - prm.y_out0) + (y - prm.y_out0)) * W + x) * 64 + c0;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = rstd * (gyv[i + j] - m1 - v[i + j] * m2);
            *reinterpret_cast<uint4*>(dst + i) = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]),
                                                            pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<dconv::kTmemCols>(tmem_base);
  }
}

namespace {
// x: NHWC bf16 tensor whose row 0 is image row `y_in0` and which holds `rows_in` rows per batch entry
template <int MODE>
int launch_decoder_conv_t(const __nv_bfloat16* x_nhwc, int rows_in, const __nv_bfloat16* w9, dconv::Params prm,
                          double flop, double bytes, cudaStream_t stream) {
  using namespace dconv;
  BSEG_REQUIRE(prm.H % kTileH == 0 && prm.W % kTileW == 0, "decoder_conv: H=%d W=%d must be multiples of %d x %d",
               prm.H, prm.W, kTileH, kTileW);
  BSEG_REQUIRE(prm.ty_begin >= 0 && prm.ty_begin * kTileH < prm.H, "decoder_conv: tile range");
  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {64, static_cast<uint64_t>(prm.W), static_cast<uint64_t>(rows_in), static_cast<uint64_t>(prm.B)};
    uint64_t strides[3] = {128, static_cast<uint64_t>(prm.W) * 128, static_cast<uint64_t>(rows_in) * prm.W * 128};
    uint32_t box[4] = {64, kTileW, kTileH + 2, 1};  // the tile's rows plus one halo row above and below
    int rc = make_tmap_bf16(&tx, x_nhwc, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    int rc = make_tmap_bf16_2d(&tw, w9, 64, 9 * 64, 64, 64, 192);
    if (rc) return rc;
  }
  auto kern = decoder_conv_kernel<MODE>;
  static PerDeviceFlag attr_once;
  if (attr_once.first()) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  }
  const long long tiles = static_cast<long long>(prm.B) * (prm.H / kTileH - prm.ty_begin) * (prm.W / kTileW);
  const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
  ProfScope prof(CAT_DECODER_HEAD, flop, bytes, stream);
  BSEG_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), kSmemBytes, stream, tx, tw, prm));
  count_launch();
  return 0;
}
}  // namespace

int launch_decoder_head(const __nv_bfloat16* x_nhwc, const __nv_bfloat16* w9, const float* conv_b, const float* ln_w,
                        const float* ln_b, const float* head_w, const float* head_b, float* pred, int B, int H, int W,
                        float eps, int y_begin, cudaStream_t stream) {
  BSEG_REQUIRE(y_begin >= 0 && y_begin < H && y_begin % dconv::kTileH == 0, "decoder_head: y_begin=%d", y_begin);
  dconv::Params prm{};
  prm.conv_b = conv_b; prm.ln_w = ln_w; prm.ln_b = ln_b; prm.head_w = head_w; prm.head_b = head_b;
  prm.pred = pred;
  prm.B = B; prm.H = H; prm.W = W; prm.eps = eps;
  prm.ty_begin = y_begin / dconv::kTileH;  // image rows < y_begin are not computed (pred is left untouched there)
  const double rows = H - y_begin;
  return launch_decoder_conv_t<dconv::kModeHead>(x_nhwc, H, w9, prm,
                                                 static_cast<double>(B) * rows * W * (2.0 * 576 * 64 + 2.0 * 64 * 3),
                                                 static_cast<double>(B) * rows * W * (64 * 2 + 3 * 4), stream);
}

// Backward through conv1x1 head, GELU and LayerNorm(C) for image rows >= y0 (the loss only touches the query half,
// src/model.py:48-57): recomputes the conv3x3 output from the saved NHWC input and writes d(conv out) as
// bf16 NHWC [B, H - y0, W, 64].
int launch_decoder_head_bwd(const __nv_bfloat16* x_nhwc, const __nv_bfloat16* w9, const float* conv_b,
                            const float* ln_w, const float* ln_b, const float* head_w, const float* head_b,
                            const float* d_pred, __nv_bfloat16* d_conv, int B, int H, int W, int y0, float eps,
                            cudaStream_t stream) {
  BSEG_REQUIRE(y0 >= 0 && y0 < H && y0 % dconv::kTileH == 0, "decoder_head_bwd: y0=%d", y0);
  dconv::Params prm{};
  prm.conv_b = conv_b; prm.ln_w = ln_w; prm.ln_b = ln_b; prm.head_w = head_w; prm.head_b = head_b;
  prm.d_pred = d_pred;
  prm.out_bf16 = d_conv;
  prm.B = B; prm.H = H; prm.W = W; prm.eps = eps;
  prm.ty_begin = y0 / dconv::kTileH;
  prm.y_out0 = y0;
  const double px = static_cast<double>(B) * (H - y0) * W;
  return launch_decoder_conv_t<dconv::kModeHeadBwd>(x_nhwc, H, w9, prm, px * (2.0 * 576 * 64 + 4.0 * 64 * 3),
                                                    px * (64 * 2 * 2 + 3 * 4), stream);
}

// conv3x3 dgrad: d_conv is bf16 NHWC [B, H - y0, W, 64] (rows above y0 are zero); w9b = flipped / transposed taps.
// Writes rows of the decoder_embed dgrad operand [B*(H/16)*(W/16), 16384] for image rows >= y_first (a multiple of 16).
int launch_decoder_conv_dgrad(const __nv_bfloat16* d_conv, const __nv_bfloat16* w9b, __nv_bfloat16* d_dec_rows, int B,
                              int H, int W, int y0, int y_first, cudaStream_t stream) {
  BSEG_REQUIRE(y_first % 16 == 0 && y_first <= y0 && y_first >= 0, "decoder_conv_dgrad: y_first=%d y0=%d", y_first, y0);
  dconv::Params prm{};
  prm.out_bf16 = d_dec_rows;
  prm.B = B; prm.H = H; prm.W = W;
  prm.ty_begin = y_first / dconv::kTileH;
  prm.y_in0 = y0;
  const double px = static_cast<double>(B) * (H - y_first) * W;
  return launch_decoder_conv_t<dconv::kModeDgrad>(d_conv, H - y0, w9b, prm, px * 2.0 * 576 * 64, px * 64 * 2 * 2, stream);
}

}  // namespace bseg
