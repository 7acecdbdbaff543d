// Tile ingest: uint16 4-band Planet Dove scene -> model-ready 448x448 tiles, bit-faithful to the reference chain
//   tif_image 4-band branch        src/util/geo_util.py:454-468   (false-colour composite, clip, scale, u8 truncate)
//   crop_tif / padded_crop         src/util/geo_util.py:297-341   (zero padding; nodata padded with 1)
//   PIL resize(BICUBIC) + /255     src/data.py:93-124             (two-pass fixed-point 8-bit resampler)
//   K.Normalize(mean, std)         src/data.py:226-229
// The scene-global statistics of tif_image (min over valid pixels, per-channel max) are a separate reduction.
#include <type_traits>

#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

// ----------------------------------------------------------------------------------------------
// scene statistics.  Composite channels: c0 = band3, c1 = band2, c2 = mean(band0, band1).
//   stats[0] = min over valid pixels and the 3 channels   (geo_util.py:459)
//   stats[1..3] = per-channel max over ALL pixels          (geo_util.py:463-464 runs before the nodata zeroing)
// Scene type T: uint16 (raw Planet Dove counts) or float32 (the reference's reprojected / averaged mosaic, which may
// contain negative values after cubic resampling).  Floats are mapped to order-preserving unsigned keys for the atomics.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int float_key(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
// pixels [first, first + count) of a scene whose bands are npix pixels apart
template <typename T>
__global__ void scene_stats_kernel(const T* __restrict__ scene, const uint8_t* __restrict__ nodata, long long npix,
                                   long long first, long long count, unsigned int* __restrict__ scratch) {
  float mn = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
  for (long long i = first + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < first + count;
       i += (long long)gridDim.x * blockDim.x) {
    const float b0 = scene[i], b1 = scene[npix + i], b2 = scene[2 * npix + i], b3 = scene[3 * npix + i];
    const float c2 = __fmul_rn(__fadd_rn(b0, b1), 0.5f);
    mx0 = fmaxf(mx0, b3);
    mx1 = fmaxf(mx1, b2);
    mx2 = fmaxf(mx2, c2);
    if (!nodata[i]) mn = fminf(mn, fminf(b3, fminf(b2, c2)));
  }
  mn = -warp_max(-mn);
  mx0 = warp_max(mx0);
  mx1 = warp_max(mx1);
  mx2 = warp_max(mx2);
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&scratch[0], float_key(mn));
    atomicMax(&scratch[1], float_key(mx0));
    atomicMax(&scratch[2], float_key(mx1));
    atomicMax(&scratch[3], float_key(mx2));
  }
}
__global__ void scene_stats_init_kernel(unsigned int* scratch) {
  scratch[0] = 0xFFFFFFFFu;                    // key of the largest float
  scratch[1] = scratch[2] = scratch[3] = 0u;   // key of the smallest
}
__global__ void scene_stats_final_kernel(const unsigned int* scratch, float* stats) {
  for (int i = 0; i < 4; ++i) stats[i] = key_float(scratch[i]);
}
namespace {
// rows [row0, row1) of the scene; stats == nullptr leaves the (order-preserving) keys in scratch for a later merge
template <typename T>
int launch_scene_stats_t(const T* scene, const uint8_t* nodata, int Hs, int Ws, int row0, int row1, float* stats,
                         unsigned int* scratch, cudaStream_t stream) {
  const long long npix = static_cast<long long>(Hs) * Ws;
  const long long first = static_cast<long long>(row0) * Ws, count = static_cast<long long>(row1 - row0) * Ws;
  scene_stats_init_kernel<<<1, 1, 0, stream>>>(scratch);
  if (count > 0) {
    long long blocks = (count + 2047) / 2048;
    if (blocks > 148 * 8) blocks = 148 * 8;
    scene_stats_kernel<T><<<static_cast<int>(blocks), 256, 0, stream>>>(scene, nodata, npix, first, count, scratch);
  }
  if (stats != nullptr) scene_stats_final_kernel<<<1, 1, 0, stream>>>(scratch, stats);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
}  // namespace
int launch_scene_stats_rows(const void* scene, int is_f32, const uint8_t* nodata, int Hs, int Ws, int row0, int row1,
                            unsigned int* scratch, cudaStream_t stream) {
  return is_f32 ? launch_scene_stats_t(static_cast<const float*>(scene), nodata, Hs, Ws, row0, row1, nullptr, scratch,
                                       stream)
                : launch_scene_stats_t(static_cast<const uint16_t*>(scene), nodata, Hs, Ws, row0, row1, nullptr,
                                       scratch, stream);
}
int launch_scene_stats_finalize(const unsigned int* scratch, float* stats, cudaStream_t stream) {
  scene_stats_final_kernel<<<1, 1, 0, stream>>>(scratch, stats);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_scene_stats(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, float* stats,
                       unsigned int* scratch, cudaStream_t stream) {
  return launch_scene_stats_t(scene, nodata, Hs, Ws, 0, Hs, stats, scratch, stream);
}
int launch_scene_stats_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, float* stats,
                           unsigned int* scratch, cudaStream_t stream) {
  return launch_scene_stats_t(scene, nodata, Hs, Ws, 0, Hs, stats, scratch, stream);
}

// ----------------------------------------------------------------------------------------------
// merge_tifs accumulation (src/util/geo_util.py:410-422): N rasters already reprojected onto the output grid
// (data fp32 [N,C,H,W], yesdata uint8 [N,H,W]) -> nodata-weighted mean fp32 [C,H,W] (0 where no raster has data) and
// the merged nodata mask uint8 [H,W] (= ~any(yesdata)).  The sum runs over n = 0..N-1 in float32 like numpy's
// axis-0 reduction, so the result is bit-identical.
// ----------------------------------------------------------------------------------------------
__global__ void merge_mosaic_kernel(const float* __restrict__ data, const uint8_t* __restrict__ yes, int N, int C,
                                    long long npix, float* __restrict__ mean, uint8_t* __restrict__ nodata) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    float wsum = 0.f;
    bool any = false;
    for (int n = 0; n < N; ++n) {
      const uint8_t y = yes[n * npix + i];
      wsum = __fadd_rn(wsum, static_cast<float>(y));
      any |= (y != 0);
    }
    for (int c = 0; c < C; ++c) {
      float acc = 0.f;
      for (int n = 0; n < N; ++n)
        acc = __fadd_rn(acc, __fmul_rn(data[((long long)n * C + c) * npix + i], static_cast<float>(yes[n * npix + i])));
      mean[c * npix + i] = wsum != 0.f ? __fdiv_rn(acc, wsum) : 0.f;
    }
    nodata[i] = any ? 0 : 1;
  }
}
int launch_merge_mosaic(const float* data, const uint8_t* yesdata, int N, int C, int Hs, int Ws, float* mean,
                        uint8_t* nodata, cudaStream_t stream) {
  const long long npix = static_cast<long long>(Hs) * Ws;
  long long blocks = (npix + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope prof(CAT_INGEST, 0, static_cast<double>(npix) * (N * (C * 4 + 1) + C * 4 + 1), stream);
  merge_mosaic_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(data, yesdata, N, C, npix, mean, nodata);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// ingest kernel: one CTA = (tile, band of `band` output rows).
//   phase 1: composite u8 of the needed crop rows -> smem (planar per channel)
//   phase 2: horizontal fixed-point pass (PIL ImagingResampleHorizontal_8bpc) -> smem
//   phase 3: vertical fixed-point pass -> u8 -> /255 -> (x-mean)/std -> outputs
// coef: [448][ksize] int32 (22-bit fixed point), bounds: [448][2] = (first input index, tap count); the same
// table serves both axes because the crop and the output are square.
// ----------------------------------------------------------------------------------------------
constexpr int kOut = 448;
// Fixed-point precision of the coefficient table (runtime): 22 = 32 - 8 - 2 for PIL (Resample.c); torchvision's uint8
// antialias kernel (the HF processor's resize, HF:image_processing_backends.py:200-251 -> ATen
// upsample_avx_bilinear_bicubic_uint8) uses int16 weights with the largest precision that keeps them below 2^15.

__device__ __forceinline__ uint8_t composite_u8(float v, float mn, float denom) {
  // img.clip(min, min+3000) - min ; img /= max ; np.array(img*255, dtype=uint8)  (float32, truncation)
  const float hi = __fadd_rn(3000.0f, mn);
  const float c = fminf(fmaxf(v, mn), hi);
  const float x = __fmul_rn(__fdiv_rn(__fsub_rn(c, mn), denom), 255.0f);
  return static_cast<uint8_t>(static_cast<int>(x));
}
__device__ __forceinline__ uint8_t clip8(int v, int prec) {
  v >>= prec;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// SRC 0 / 2: uint16 / float32 4-band scene + tile boxes (src/predict.py path, PIL table, (u/255 - mean)/std)
// SRC 1     : uint8 RGB crops [n,crop,crop,3] (HWC) or [n,3,crop,crop] (CHW) (src/predict_no_prompt.py path:
//             SegGptImageProcessor.preprocess, torchvision table, (u - 255 mean)/(255 std))
enum : int { kSrcU16 = 0, kSrcU8 = 1, kSrcF32 = 2 };
template <int SRC>
__global__ void __launch_bounds__(256)
ingest_kernel(const void* __restrict__ scene_v, const uint8_t* __restrict__ nodata, int Hs, int Ws,
              const float* __restrict__ stats, const int* __restrict__ boxes, int crop, const int* __restrict__ coef,
              const int* __restrict__ bounds, int ksize, int band, int max_rows, float m0, float m1, float m2,
              float s0, float s1, float s2, float* __restrict__ out_nchw, __nv_bfloat16* __restrict__ out_patch,
              long long patch_tile_stride, uint8_t* __restrict__ out_u8, uint8_t* __restrict__ out_nodata,
              const uint8_t* __restrict__ src_u8, int src_chw, int prec) {
  constexpr bool kFromU8 = (SRC == kSrcU8);
  using SceneT = typename std::conditional<SRC == kSrcF32, float, uint16_t>::type;
  const SceneT* __restrict__ scene = static_cast<const SceneT*>(scene_v);
  extern __shared__ uint8_t sm[];
  uint8_t* comp = sm;                                   // [3][max_rows][crop]
  uint8_t* hbuf = sm + 3 * max_rows * crop;             // [3][max_rows][kOut]
  const int tile = blockIdx.y;
  const int oy0 = blockIdx.x * band;
  const int oy1 = min(oy0 + band, kOut);
  const int r0 = bounds[2 * oy0];                                      // first crop row needed
  const int r1 = bounds[2 * (oy1 - 1)] + bounds[2 * (oy1 - 1) + 1];    // one past the last
  const int nrows = r1 - r0;
  const long long npix = static_cast<long long>(Hs) * Ws;

  if constexpr (kFromU8) {
    // ---- phase 1: copy the needed rows of the uint8 crop ----
    const long long plane = static_cast<long long>(crop) * crop;
    const uint8_t* img = src_u8 + static_cast<long long>(tile) * 3 * plane;
    for (int i = threadIdx.x; i < nrows * crop; i += blockDim.x) {
      const int rr = i / crop, cx = i % crop;
      const long long p = static_cast<long long>(r0 + rr) * crop + cx;
#pragma unroll
      for (int c = 0; c < 3; ++c) comp[(c * max_rows + rr) * crop + cx] = src_chw ? img[c * plane + p] : img[p * 3 + c];
    }
  } else {
  const int xmin = boxes[tile * 4 + 0], ymin = boxes[tile * 4 + 1];
  const float mn = stats[0];
  const float hi = __fadd_rn(3000.0f, mn);
  float den[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) den[c] = __fsub_rn(fminf(fmaxf(stats[1 + c], mn), hi), mn);

  // ---- phase 1: composite ----
  for (int i = threadIdx.x; i < nrows * crop; i += blockDim.x) {
    const int rr = i / crop, cx = i % crop;
    const int sy = ymin + r0 + rr, sx = xmin + cx;
    uint8_t v0 = 0, v1 = 0, v2 = 0, nd = 1;
    if (sy >= 0 && sy < Hs && sx >= 0 && sx < Ws) {
      const long long p = static_cast<long long>(sy) * Ws + sx;
      nd = nodata[p] ? 1 : 0;
      if (!nd) {
        const float b0 = scene[p], b1 = scene[npix + p], b2 = scene[2 * npix + p], b3 = scene[3 * npix + p];
        v0 = composite_u8(b3, mn, den[0]);
        v1 = composite_u8(b2, mn, den[1]);
        v2 = composite_u8(__fmul_rn(__fadd_rn(b0, b1), 0.5f), mn, den[2]);
      }
    }
    comp[(0 * max_rows + rr) * crop + cx] = v0;
    comp[(1 * max_rows + rr) * crop + cx] = v1;
    comp[(2 * max_rows + rr) * crop + cx] = v2;
    // optional crop-resolution outputs; each crop row is owned by the band whose first needed row is <= it
    // (bands overlap, identical values are written, which is benign)
    if (out_u8) {
      uint8_t* o = out_u8 + ((static_cast<long long>(tile) * crop + (r0 + rr)) * crop + cx) * 3;
      o[0] = v0; o[1] = v1; o[2] = v2;
    }
    if (out_nodata) out_nodata[(static_cast<long long>(tile) * crop + (r0 + rr)) * crop + cx] = nd;
  }
  }
  __syncthreads();

  // ---- phase 2: horizontal pass ----
  for (int i = threadIdx.x; i < nrows * kOut; i += blockDim.x) {
    const int rr = i / kOut, ox = i % kOut;
    const int x0 = bounds[2 * ox], cnt = bounds[2 * ox + 1];
    const int* k = coef + ox * ksize;
    int a0 = 1 << (prec - 1), a1 = a0, a2 = a0;
    const uint8_t* c0 = comp + (0 * max_rows + rr) * crop + x0;
    const uint8_t* c1 = comp + (1 * max_rows + rr) * crop + x0;
    const uint8_t* c2 = comp + (2 * max_rows + rr) * crop + x0;
    for (int t = 0; t < cnt; ++t) {
      const int w = __ldg(k + t);
      a0 += c0[t] * w;
      a1 += c1[t] * w;
      a2 += c2[t] * w;
    }
    hbuf[(0 * max_rows + rr) * kOut + ox] = clip8(a0, prec);
    hbuf[(1 * max_rows + rr) * kOut + ox] = clip8(a1, prec);
    hbuf[(2 * max_rows + rr) * kOut + ox] = clip8(a2, prec);
  }
  __syncthreads();

  // ---- phase 3: vertical pass + normalise ----
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  const int nout = (oy1 - oy0) * kOut;
  for (int i = threadIdx.x; i < 3 * nout; i += blockDim.x) {
    const int c = i / nout;
    const int oy = oy0 + (i % nout) / kOut, ox = i % kOut;
    const int y0 = bounds[2 * oy] - r0, cnt = bounds[2 * oy + 1];
    const int* k = coef + oy * ksize;
    int a = 1 << (prec - 1);
    const uint8_t* h = hbuf + (c * max_rows + y0) * kOut + ox;
    for (int t = 0; t < cnt; ++t) a += h[t * kOut] * __ldg(k + t);
    const float u = static_cast<float>(clip8(a, prec));
    // PIL path: (u/255 - mean)/std (src/data.py:101,226-229); HF processor: (u - 255 mean)/(255 std), mean/std passed
    // pre-multiplied (HF:image_processing_backends.py:292-331)
    const float val = kFromU8 ? __fdiv_rn(__fsub_rn(u, mean[c]), stdv[c])
                              : __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), mean[c]), stdv[c]);
    if (out_nchw) out_nchw[((static_cast<long long>(tile) * 3 + c) * kOut + oy) * kOut + ox] = val;
    if (out_patch) {
      // row of the patch-embedding A operand: token (oy/16, ox/16), k = c*256 + (oy%16)*16 + ox%16
      const long long row = static_cast<long long>(oy >> 4) * 28 + (ox >> 4);
      out_patch[tile * patch_tile_stride + row * 768 + c * 256 + (oy & 15) * 16 + (ox & 15)] =
          __float2bfloat16_rn(val);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// ingest kernel v2 (used whenever the resampling table has <= 12 taps): same three phases, restructured so that one
// shared-memory word feeds four multiply-adds.
//   * the composite is stored TRANSPOSED AND PACKED: word (ch, x) of a row group holds the bytes of 4 consecutive crop
//     rows, so a horizontal tap is one LDS.32 + 4 byte extracts + 4 IMADs (thread == output column, its <= KS
//     coefficients live in registers); row groups stream through a double buffer, one __syncthreads per group;
//   * the horizontal result is stored planar with x contiguous, so a vertical tap is again one LDS.32 for 4 outputs
//     (thread == 4 adjacent output columns of one output row and channel);
//   * u8 -> normalised float goes through a 3 x 256 table built per CTA with the exact IEEE divisions of the
//     reference ((u/255 - mean)/std resp. (u - 255 mean)/(255 std)), so the epilogue is a lookup.
// Thread x of the composite phase owns crop column x of all 4 rows of the group: global loads are coalesced along x
// and the packed word is written with one conflict-free STS.32 per channel.
// ----------------------------------------------------------------------------------------------
constexpr int kV2Threads = 512;
template <int SRC, int KS>
__global__ void __launch_bounds__(kV2Threads, 2)
ingest_v2_kernel(const void* __restrict__ scene_v, const uint8_t* __restrict__ nodata, int Hs, int Ws,
                 const float* __restrict__ stats, const int* __restrict__ boxes, int crop, const int* __restrict__ coef,
                 const int* __restrict__ bounds, int ksize, int band, int rows_pad, float m0, float m1, float m2,
                 float s0, float s1, float s2, float* __restrict__ out_nchw, __nv_bfloat16* __restrict__ out_patch,
                 long long patch_tile_stride, uint8_t* __restrict__ out_u8, uint8_t* __restrict__ out_nodata,
                 const uint8_t* __restrict__ src_u8, int src_chw, int prec) {
  constexpr bool kFromU8 = (SRC == kSrcU8);
  using SceneT = typename std::conditional<SRC == kSrcF32, float, uint16_t>::type;
  const SceneT* __restrict__ scene = static_cast<const SceneT*>(scene_v);
  extern __shared__ __align__(16) uint8_t sm[];
  uint32_t* comp = reinterpret_cast<uint32_t*>(sm);                 // [2][3][crop] packed 4-row words
  uint8_t* hbuf = sm + 2 * 3 * crop * 4;                            // [3][rows_pad][kOut]
  float* lut = reinterpret_cast<float*>(hbuf + 3 * rows_pad * kOut);  // [3][256]
  const int tid = threadIdx.x;
  const int tile = blockIdx.y;
  const int oy0 = blockIdx.x * band;
  const int oy1 = min(oy0 + band, kOut);
  const int r0 = bounds[2 * oy0];
  const int r1 = bounds[2 * (oy1 - 1)] + bounds[2 * (oy1 - 1) + 1];
  const int nrows = r1 - r0;
  const int ngroups = (nrows + 3) >> 2;
  const long long npix = static_cast<long long>(Hs) * Ws;

  {
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    for (int i = tid; i < 768; i += kV2Threads) {
      const int c = i >> 8;
      const float u = static_cast<float>(i & 255);
      lut[i] = kFromU8 ? __fdiv_rn(__fsub_rn(u, mean[c]), stdv[c])
                       : __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), mean[c]), stdv[c]);
    }
  }
  // horizontal pass: this thread's output column and its coefficients
  int kw[KS];
  int hx0 = 0;
  if (tid < kOut) {
    hx0 = bounds[2 * tid];
    const int hcnt = bounds[2 * tid + 1];
#pragma unroll
    for (int t = 0; t < KS; ++t) kw[t] = (t < hcnt && t < ksize) ? coef[tid * ksize + t] : 0;
  } else {
#pragma unroll
    for (int t = 0; t < KS; ++t) kw[t] = 0;
  }

  int xmin = 0, ymin = 0;
  float mn = 0.f, den[3] = {1.f, 1.f, 1.f};
  if constexpr (!kFromU8) {
    xmin = boxes[tile * 4 + 0];
    ymin = boxes[tile * 4 + 1];
    mn = stats[0];
    const float hi = __fadd_rn(3000.0f, mn);
#pragma unroll
    for (int c = 0; c < 3; ++c) den[c] = __fsub_rn(fminf(fmaxf(stats[1 + c], mn), hi), mn);
  }
  const long long plane = static_cast<long long>(crop) * crop;

  auto composite = [&](int g, int buf) {
    for (int x = tid; x < crop; x += kV2Threads) {
      uint32_t w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rr = 4 * g + r;
        if (rr >= nrows) break;
        uint32_t v0 = 0, v1 = 0, v2 = 0;
        if constexpr (kFromU8) {
          const uint8_t* img = src_u8 + static_cast<long long>(tile) * 3 * plane;
          const long long p = static_cast<long long>(r0 + rr) * crop + x;
          if (src_chw) { v0 = img[p]; v1 = img[plane + p]; v2 = img[2 * plane + p]; }
          else { v0 = img[p * 3]; v1 = img[p * 3 + 1]; v2 = img[p * 3 + 2]; }
        } else {
          const int sy = ymin + r0 + rr, sx = xmin + x;
          uint8_t nd = 1;
          if (sy >= 0 && sy < Hs && sx >= 0 && sx < Ws) {
            const long long p = static_cast<long long>(sy) * Ws + sx;
            nd = nodata[p] ? 1 : 0;
            if (!nd) {
              const float b0 = scene[p], b1 = scene[npix + p], b2 = scene[2 * npix + p], b3 = scene[3 * npix + p];
              v0 = composite_u8(b3, mn, den[0]);
              v1 = composite_u8(b2, mn, den[1]);
              v2 = composite_u8(__fmul_rn(__fadd_rn(b0, b1), 0.5f), mn, den[2]);
            }
          }
          // crop-resolution outputs; neighbouring bands write identical values to the rows they share
          if (out_u8) {
            uint8_t* o = out_u8 + ((static_cast<long long>(tile) * crop + (r0 + rr)) * crop + x) * 3;
            o[0] = static_cast<uint8_t>(v0); o[1] = static_cast<uint8_t>(v1); o[2] = static_cast<uint8_t>(v2);
          }
          if (out_nodata) out_nodata[(static_cast<long long>(tile) * crop + (r0 + rr)) * crop + x] = nd;
        }
        w0 |= v0 << (8 * r);
        w1 |= v1 << (8 * r);
        w2 |= v2 << (8 * r);
      }
      comp[(buf * 3 + 0) * crop + x] = w0;
      comp[(buf * 3 + 1) * crop + x] = w1;
      comp[(buf * 3 + 2) * crop + x] = w2;
    }
  };
  auto horizontal = [&](int g, int buf) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const uint32_t* cp = comp + (buf * 3 + ch) * crop;
      int a0 = 1 << (prec - 1), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
      for (int t = 0; t < KS; ++t) {
        const uint32_t w = cp[min(hx0 + t, crop - 1)];
        const int k = kw[t];
        a0 += static_cast<int>(w & 0xFFu) * k;
        a1 += static_cast<int>((w >> 8) & 0xFFu) * k;
        a2 += static_cast<int>((w >> 16) & 0xFFu) * k;
        a3 += static_cast<int>(w >> 24) * k;
      }
      uint8_t* hp = hbuf + (ch * rows_pad + 4 * g) * kOut + tid;
      hp[0] = clip8(a0, prec);
      hp[kOut] = clip8(a1, prec);
      hp[2 * kOut] = clip8(a2, prec);
      hp[3 * kOut] = clip8(a3, prec);
    }
  };

  composite(0, 0);
  for (int g = 0; g < ngroups; ++g) {
    __syncthreads();
    if (g + 1 < ngroups) composite(g + 1, (g + 1) & 1);
    if (tid < kOut) horizontal(g, g & 1);
  }
  __syncthreads();

  // ---- vertical pass + normalise: item = (output row, channel, 4 adjacent output columns) ----
  const int items = (oy1 - oy0) * 3 * (kOut / 4);
  for (int i = tid; i < items; i += kV2Threads) {
    const int q = i % (kOut / 4);
    const int c = (i / (kOut / 4)) % 3;
    const int oy = oy0 + i / (3 * (kOut / 4));
    const int y0 = bounds[2 * oy] - r0, cnt = bounds[2 * oy + 1];
    const int* k = coef + oy * ksize;
    int a0 = 1 << (prec - 1), a1 = a0, a2 = a0, a3 = a0;
    const uint8_t* hp = hbuf + (c * rows_pad + y0) * kOut + 4 * q;
    for (int t = 0; t < cnt; ++t) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(hp + t * kOut);
      const int kk = __ldg(k + t);
      a0 += static_cast<int>(w & 0xFFu) * kk;
      a1 += static_cast<int>((w >> 8) & 0xFFu) * kk;
      a2 += static_cast<int>((w >> 16) & 0xFFu) * kk;
      a3 += static_cast<int>(w >> 24) * kk;
    }
    const float* lc = lut + c * 256;
    const float f0 = lc[clip8(a0, prec)], f1 = lc[clip8(a1, prec)], f2 = lc[clip8(a2, prec)], f3 = lc[clip8(a3, prec)];
    const int ox = 4 * q;
    if (out_nchw)
      *reinterpret_cast<float4*>(out_nchw + ((static_cast<long long>(tile) * 3 + c) * kOut + oy) * kOut + ox) =
          make_float4(f0, f1, f2, f3);
    if (out_patch) {
      const long long row = static_cast<long long>(oy >> 4) * 28 + (ox >> 4);
      *reinterpret_cast<uint2*>(out_patch + tile * patch_tile_stride + row * 768 + c * 256 + (oy & 15) * 16 + (ox & 15)) =
          make_uint2(pack_bf16x2(f0, f1), pack_bf16x2(f2, f3));
    }
  }
}

namespace {
// rows of output per CTA for the v2 kernel: the largest band whose shared memory lets two CTAs share an SM, else the
// largest that fits at all
bool ingest_v2_geometry(int crop, int* band_out, int* rows_pad_out, size_t* smem_out) {
  const double scale = static_cast<double>(crop) / 448.0;
  const double support = 2.0 * (scale < 1.0 ? 1.0 : scale);
  const int bands[] = {28, 16, 8, 4, 2, 1};
  for (int pass = 0; pass < 2; ++pass) {
    const size_t limit = pass == 0 ? 110 * 1024 : 200 * 1024;
    for (int band : bands) {
      const int rows = (static_cast<int>(band * scale + 2 * support + 4) + 3) / 4 * 4;
      const size_t smem = 24ull * crop + 3ull * rows * 448 + 3072;
      if (smem <= limit) {
        *band_out = band;
        *rows_pad_out = rows;
        *smem_out = smem;
        return true;
      }
    }
  }
  return false;
}

template <int SRC, int KS>
int launch_ingest_v2(const void* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats, const int* boxes,
                     int n_tiles, int crop, const int* coef, const int* bounds, int ksize, const float* mean,
                     const float* stdv, float* out_nchw, __nv_bfloat16* out_patch, long long patch_tile_stride,
                     uint8_t* out_u8, uint8_t* out_nodata, const uint8_t* src_u8, int src_chw, int prec,
                     cudaStream_t stream) {
  int band, rows_pad;
  size_t smem;
  BSEG_REQUIRE(ingest_v2_geometry(crop, &band, &rows_pad, &smem), "ingest: crop=%d does not fit in shared memory", crop);
  auto kern = ingest_v2_kernel<SRC, KS>;
  static PerDeviceInt attr_cache;  // largest opt-in size set so far, per template instantiation and device
  int& attr_bytes = attr_cache.get();
  if (static_cast<int>(smem) > attr_bytes) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_bytes = static_cast<int>(smem);
  }
  dim3 grid((kOut + band - 1) / band, n_tiles);
  constexpr bool kFromU8 = (SRC == kSrcU8);
  const double in_bytes = kFromU8 ? 3.0 : (SRC == kSrcF32 ? 17.0 : 9.0);  // per crop pixel: bands + nodata
  ProfScope prof(CAT_INGEST, 0,
                 static_cast<double>(n_tiles) * (in_bytes * crop * crop + 3.0 * 448 * 448 * ((out_nchw ? 4 : 0) + (out_patch ? 2 : 0))),
                 stream);
  kern<<<grid, kV2Threads, smem, stream>>>(scene, nodata, Hs, Ws, stats, boxes, crop, coef, bounds, ksize, band,
                                           rows_pad, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2], out_nchw,
                                           out_patch, patch_tile_stride, out_u8, out_nodata, src_u8, src_chw, prec);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
}  // namespace

namespace {
template <int SRC>
int launch_ingest_t(const void* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats, const int* boxes,
                    int n_tiles, int crop, const int* coef, const int* bounds, int ksize, int band, int max_rows,
                    const float* mean, const float* stdv, float* out_nchw, __nv_bfloat16* out_patch,
                    long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, const uint8_t* src_u8,
                    int src_chw, int prec, cudaStream_t stream) {
  if (n_tiles == 0) return 0;
  BSEG_REQUIRE(band > 0 && max_rows > 0 && crop > 0, "ingest: bad geometry");
  BSEG_REQUIRE(prec >= 1 && prec <= 22, "ingest: coefficient precision %d out of range", prec);
  if (ksize <= 8)
    return launch_ingest_v2<SRC, 8>(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, coef, bounds, ksize, mean, stdv,
                                    out_nchw, out_patch, patch_tile_stride, out_u8, out_nodata, src_u8, src_chw, prec,
                                    stream);
  if (ksize <= 12)
    return launch_ingest_v2<SRC, 12>(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, coef, bounds, ksize, mean,
                                     stdv, out_nchw, out_patch, patch_tile_stride, out_u8, out_nodata, src_u8, src_chw,
                                     prec, stream);
  const size_t smem = static_cast<size_t>(3) * max_rows * (crop + kOut);
  BSEG_REQUIRE(smem <= 200 * 1024, "ingest: crop=%d band=%d needs %zu B of shared memory", crop, band, smem);
  constexpr bool kFromU8 = (SRC == kSrcU8);
  auto kern = ingest_kernel<SRC>;
  static PerDeviceInt attr_cache;  // largest opt-in size set so far, per template instantiation and device
  int& attr_bytes = attr_cache.get();
  if (static_cast<int>(smem) > attr_bytes) {
    BSEG_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_bytes = static_cast<int>(smem);
  }
  dim3 grid((kOut + band - 1) / band, n_tiles);
  ProfScope prof(CAT_INGEST, 0,
                 static_cast<double>(n_tiles) * ((kFromU8 ? 3.0 : 8.0) * crop * crop + 3.0 * 448 * 448 * (out_nchw ? 4 : 2)),
                 stream);
  kern<<<grid, 256, smem, stream>>>(scene, nodata, Hs, Ws, stats, boxes, crop, coef, bounds, ksize, band, max_rows,
                                    mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2], out_nchw, out_patch,
                                    patch_tile_stride, out_u8, out_nodata, src_u8, src_chw, prec);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
}  // namespace

int launch_ingest(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats, const int* boxes,
                  int n_tiles, int crop, const int* coef, const int* bounds, int ksize, int band, int max_rows,
                  const float* mean, const float* stdv, float* out_nchw, __nv_bfloat16* out_patch,
                  long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream) {
  return launch_ingest_t<kSrcU16>(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, coef, bounds, ksize, band,
                                  max_rows, mean, stdv, out_nchw, out_patch, patch_tile_stride, out_u8, out_nodata,
                                  nullptr, 0, 22, stream);
}

int launch_ingest_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats, const int* boxes,
                      int n_tiles, int crop, const int* coef, const int* bounds, int ksize, int band, int max_rows,
                      const float* mean, const float* stdv, float* out_nchw, __nv_bfloat16* out_patch,
                      long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream) {
  return launch_ingest_t<kSrcF32>(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, coef, bounds, ksize, band,
                                  max_rows, mean, stdv, out_nchw, out_patch, patch_tile_stride, out_u8, out_nodata,
                                  nullptr, 0, 22, stream);
}

int launch_preprocess_u8(const uint8_t* images, int chw, int n, int crop, const int* coef, const int* bounds, int ksize,
                         int prec, int band, int max_rows, const float* mean255, const float* std255, float* out_nchw,
                         cudaStream_t stream) {
  return launch_ingest_t<kSrcU8>(nullptr, nullptr, 0, 0, nullptr, nullptr, n, crop, coef, bounds, ksize, band, max_rows,
                               mean255, std255, out_nchw, nullptr, 0, nullptr, nullptr, images, chw, prec, stream);
}

// ----------------------------------------------------------------------------------------------
// Native-resolution ingest: the model runs at the tile size (image_size == crop), so
// get_crop skips the PIL resize (src/data.py:94, `if inpt_size != crop_size`) and the chain is purely per pixel:
// composite -> u8 -> /255 -> (x - mean)/std, same arithmetic as above (the normalise goes through a 3 x 256 table of
// the exact IEEE results).  HBM-bound: 9 B in (4 x u16 + nodata), 12 B out (+3 / +1 for the optional u8 / nodata
// crops) per pixel; one thread handles eight consecutive pixels with 16-byte band loads and stores when the box is
// aligned and inside the scene, else pixel by pixel with the reference's zero / nodata padding.
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024)
ingest_native_kernel(const T* __restrict__ scene, const uint8_t* __restrict__ nodata, int Hs, int Ws,
                     const float* __restrict__ stats, const int* __restrict__ boxes, int crop, float m0, float m1,
                     float m2, float s0, float s1, float s2, float* __restrict__ out_nchw, uint8_t* __restrict__ out_u8,
                     uint8_t* __restrict__ out_nodata) {
  __shared__ float lut[3][256];
  // uint16 scenes: the clipped, min-subtracted composite takes 3001 values (channels 0, 1) / 6001 half-integer values
  // (channel 2 = mean of two bands), so the three IEEE divisions per pixel become table look-ups of the exact results
  __shared__ uint8_t clut[3001 + 3001 + 6001 + 1];
  const float mn = stats[0];
  const float hi = __fadd_rn(3000.0f, mn);
  float den[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) den[c] = __fsub_rn(fminf(fmaxf(stats[1 + c], mn), hi), mn);
  const bool use_clut = sizeof(T) == 2 && mn >= 0.f && mn <= 65535.f && mn == floorf(mn);
  {
    const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
      const int c = i >> 8, u = i & 255;
      lut[c][u] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(u), 255.0f), mean[c]), stdv[c]);
    }
    if (use_clut) {
      for (int i = threadIdx.x; i < 3001 + 3001 + 6001; i += blockDim.x) {
        const int c = i < 3001 ? 0 : (i < 6002 ? 1 : 2);
        const float v = c == 0 ? mn + static_cast<float>(i) : c == 1 ? mn + static_cast<float>(i - 3001)
                                                                      : mn + 0.5f * static_cast<float>(i - 6002);
        clut[i] = composite_u8(v, mn, den[c]);
      }
    }
  }
  __syncthreads();
  const int mn_i = static_cast<int>(mn);
  const int tile = blockIdx.y;
  const int xmin = boxes[tile * 4 + 0], ymin = boxes[tile * 4 + 1];
  const long long npix = static_cast<long long>(Hs) * Ws;
  const long long plane = static_cast<long long>(crop) * crop;
  float* o_tile = out_nchw + static_cast<long long>(tile) * 3 * plane;
  // eight consecutive pixels per thread and iteration: 16-byte band loads, 8-byte nodata load, 2 x 16-byte stores per
  // channel (enough bytes in flight per SM to cover the HBM latency with one 1024-thread CTA)
  constexpr int PX = 8;
  const bool fast = sizeof(T) == 2 && (crop % PX) == 0 && (xmin % PX) == 0 && (Ws % PX) == 0 && xmin >= 0 && ymin >= 0 &&
                    xmin + crop <= Ws && ymin + crop <= Hs;
  const int groups = (crop + PX - 1) / PX;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < crop * groups; i += gridDim.x * blockDim.x) {
    const int row = i / groups, qx = (i % groups) * PX;
    uint8_t v[PX][3], nd[PX];
    if (fast) {
      const long long p = static_cast<long long>(ymin + row) * Ws + xmin + qx;
      const uint2 ndw = *reinterpret_cast<const uint2*>(nodata + p);
      uint4 b[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) b[k] = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(scene) + k * npix + p);
      const unsigned short* b0 = reinterpret_cast<const unsigned short*>(&b[0]);
      const unsigned short* b1 = reinterpret_cast<const unsigned short*>(&b[1]);
      const unsigned short* b2 = reinterpret_cast<const unsigned short*>(&b[2]);
      const unsigned short* b3 = reinterpret_cast<const unsigned short*>(&b[3]);
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        nd[j] = (((j < 4 ? ndw.x : ndw.y) >> (8 * (j & 3))) & 0xffu) ? 1 : 0;
        v[j][0] = v[j][1] = v[j][2] = 0;
        if (!nd[j]) {
          if (use_clut) {
            v[j][0] = clut[min(max(static_cast<int>(b3[j]) - mn_i, 0), 3000)];
            v[j][1] = clut[3001 + min(max(static_cast<int>(b2[j]) - mn_i, 0), 3000)];
            v[j][2] = clut[6002 + min(max(static_cast<int>(b0[j]) + static_cast<int>(b1[j]) - 2 * mn_i, 0), 6000)];
          } else {
            v[j][0] = composite_u8(static_cast<float>(b3[j]), mn, den[0]);
            v[j][1] = composite_u8(static_cast<float>(b2[j]), mn, den[1]);
            v[j][2] = composite_u8(__fmul_rn(__fadd_rn(static_cast<float>(b0[j]), static_cast<float>(b1[j])), 0.5f), mn, den[2]);
          }
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        const int sy = ymin + row, sx = xmin + qx + j;
        v[j][0] = v[j][1] = v[j][2] = 0;
        nd[j] = 1;
        if (qx + j < crop && sy >= 0 && sy < Hs && sx >= 0 && sx < Ws) {
          const long long p = static_cast<long long>(sy) * Ws + sx;
          nd[j] = nodata[p] ? 1 : 0;
          if (!nd[j]) {
            const float b0 = scene[p], b1 = scene[npix + p], b2 = scene[2 * npix + p], b3 = scene[3 * npix + p];
            v[j][0] = composite_u8(b3, mn, den[0]);
            v[j][1] = composite_u8(b2, mn, den[1]);
            v[j][2] = composite_u8(__fmul_rn(__fadd_rn(b0, b1), 0.5f), mn, den[2]);
          }
        }
      }
    }
    const long long o = static_cast<long long>(row) * crop + qx;
    if (qx + PX <= crop && (crop & 3) == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        *reinterpret_cast<float4*>(o_tile + c * plane + o) =
            make_float4(lut[c][v[0][c]], lut[c][v[1][c]], lut[c][v[2][c]], lut[c][v[3][c]]);
        *reinterpret_cast<float4*>(o_tile + c * plane + o + 4) =
            make_float4(lut[c][v[4][c]], lut[c][v[5][c]], lut[c][v[6][c]], lut[c][v[7][c]]);
      }
    } else {
      for (int j = 0; j < PX && qx + j < crop; ++j)
#pragma unroll
        for (int c = 0; c < 3; ++c) o_tile[c * plane + o + j] = lut[c][v[j][c]];
    }
    if (out_u8 || out_nodata) {
      for (int j = 0; j < PX && qx + j < crop; ++j) {
        if (out_u8) {
          uint8_t* u = out_u8 + (static_cast<long long>(tile) * plane + o + j) * 3;
          u[0] = v[j][0]; u[1] = v[j][1]; u[2] = v[j][2];
        }
        if (out_nodata) out_nodata[static_cast<long long>(tile) * plane + o + j] = nd[j];
      }
    }
  }
}

template <typename T>
static int launch_ingest_native_t(const T* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                                  const int* boxes, int n_tiles, int crop, const float* mean, const float* stdv,
                                  float* out_nchw, uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream) {
  if (n_tiles == 0) return 0;
  BSEG_REQUIRE(crop > 0 && out_nchw != nullptr, "ingest_native: bad arguments");
  BSEG_REQUIRE((reinterpret_cast<uintptr_t>(scene) & 15) == 0 && (reinterpret_cast<uintptr_t>(nodata) & 7) == 0 &&
                   (reinterpret_cast<uintptr_t>(out_nchw) & 15) == 0,
               "ingest_native: misaligned buffers");
  // two CTAs of 1024 threads per SM over all tiles: large CTAs amortise the 12 K-entry composite table each one builds
  const int work = crop * ((crop + 7) / 8);
  int bx = (work + 1023) / 1024;
  const int cap = (num_sms() * 2 + n_tiles - 1) / n_tiles;
  if (bx > cap) bx = cap < 1 ? 1 : cap;
  ProfScope prof(CAT_INGEST, 0,
                 static_cast<double>(n_tiles) * crop * crop * ((sizeof(T) == 2 ? 9.0 : 17.0) + 12.0 + (out_u8 ? 3 : 0) +
                                                               (out_nodata ? 1 : 0)),
                 stream);
  ingest_native_kernel<T><<<dim3(bx, n_tiles), 1024, 0, stream>>>(scene, nodata, Hs, Ws, stats, boxes, crop, mean[0], mean[1],
                                                                 mean[2], stdv[0], stdv[1], stdv[2], out_nchw, out_u8,
                                                                 out_nodata);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int launch_ingest_native(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                         const int* boxes, int n_tiles, int crop, const float* mean, const float* stdv, float* out_nchw,
                         uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream) {
  return launch_ingest_native_t(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, mean, stdv, out_nchw, out_u8,
                                out_nodata, stream);
}
int launch_ingest_native_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                             const int* boxes, int n_tiles, int crop, const float* mean, const float* stdv,
                             float* out_nchw, uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream) {
  return launch_ingest_native_t(scene, nodata, Hs, Ws, stats, boxes, n_tiles, crop, mean, stdv, out_nchw, out_u8,
                                out_nodata, stream);
}

}  // namespace bseg
