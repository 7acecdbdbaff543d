// Bandwidth-bound kernels of the SegGPT tile path: LayerNorm, patchify (im2col for the stride-16 patch
// embedding), two-stream merge, feature-ensemble mean, prompt-mask colourise, palette decode, vote
// stitching, smooth-L1 loss.  Each kernel cites the reference lines it reproduces.
#include "../../include/bseg.h"
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"

namespace bseg {

static inline int blocks_for(long long n, int per_block, int cap = 148 * 16) {
  long long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return static_cast<int>(b);
}

// ----------------------------------------------------------------------------------------------
// fp32 -> bf16 (weight packing at bseg_create time)
// ----------------------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = *reinterpret_cast<const float4*>(src + i);
    uint2 o = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(dst + i) = o;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long long j = n & ~3LL; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
  }
}
int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t stream) {
  BSEG_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0,
               "f32_to_bf16: misaligned");
  f32_to_bf16_kernel<<<blocks_for(n, 1024), 256, 0, stream>>>(src, dst, n);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// LayerNorm over D=1024, fp32 in -> bf16 out.  nn.LayerNorm(eps=1e-6): modeling_seggpt.py:403-404,450.
// One warp per row; the row lives in registers (two-pass mean / variance, fp32).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm1024_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, __nv_bfloat16* __restrict__ out, long long ldo, long long M,
                     float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long row = warp_global; row < M; row += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * ldx);
    float4 v[8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = xr[lane + 32 * i];
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(s) * (1.0f / 1024.0f);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / 1024.0f) + eps);
    __nv_bfloat16* orow = out + row * ldo;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int col = (lane + 32 * i) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
      uint2 o = make_uint2(pack_bf16x2((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y),
                           pack_bf16x2((v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w));
      *reinterpret_cast<uint2*>(orow + col) = o;
    }
  }
}
int launch_layernorm1024(const float* x, long long ldx, const float* gamma, const float* beta, __nv_bfloat16* out,
                         long long ldo, long long M, float eps, cudaStream_t stream) {
  BSEG_REQUIRE(ldx % 4 == 0 && ldo % 4 == 0, "layernorm: leading dims must be multiples of 4");
  ProfScope prof(CAT_LAYERNORM, 0, static_cast<double>(M) * 1024 * 6, stream);
  BSEG_CHECK_CUDA(launch_pdl(layernorm1024_kernel, dim3(blocks_for(M, 8, 148 * 8)), dim3(256), 0, stream, x, ldx, gamma, beta,
                             out, ldo, M, eps));
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// patchify: builds the A operand of the patch-embedding GEMM (Conv2d k16 s16 == GEMM over im2col rows).
//   stream 0 rows = cat(prompt_pixel_values, pixel_values) on H      (modeling_seggpt.py:713)
//   stream 1 rows = cat(prompt_masks, prompt_masks|labels) on H      (modeling_seggpt.py:714-718); its bottom
//   half is replaced by the mask token (bool_masked_pos default, :910-917 and :177-178) so those rows are zero
//   here and the mask token is folded into the additive table of the GEMM epilogue.
//   A[(s*B + b)*1568 + ph*28 + pw][c*256 + py*16 + px]  (k order == Conv2d weight [out, c, py, px] flattened)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ px, const float* __restrict__ prompt_px,
                const float* __restrict__ prompt_mask, __nv_bfloat16* __restrict__ A, int B, int img) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread = 8 consecutive pixels of one image row of one channel; img = 448 (56 x 28 tokens) or 512 (64 x 32)
  const int gw = img >> 4, w8 = img >> 3, T = 2 * gw * gw;
  const long long total = 2LL * B * 3 * (2 * img) * w8;  // (stream, b, c, y, x8)
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x8 = static_cast<int>(idx % w8);
    long long r = idx / w8;
    const int y = static_cast<int>(r % (2 * img));
    r /= 2 * img;
    const int c = static_cast<int>(r % 3);
    r /= 3;
    const int b = static_cast<int>(r % B);
    const int s = static_cast<int>(r / B);
    const int ph = y >> 4, py = y & 15;
    const int x = x8 * 8;
    const int pw = x >> 4, pxo = x & 15;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    const bool top = y < img;
    if (s == 0 || top) {
      const float* src = (s == 0) ? (top ? prompt_px : px) : prompt_mask;
      const float* p = src + (((long long)b * 3 + c) * img + (top ? y : y - img)) * img + x;
      const float4 v0 = *reinterpret_cast<const float4*>(p);
      const float4 v1 = *reinterpret_cast<const float4*>(p + 4);
      o = make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y),
                     pack_bf16x2(v1.z, v1.w));
    }
    const long long row = ((long long)s * B + b) * T + ph * gw + pw;
    *reinterpret_cast<uint4*>(A + row * 768 + c * 256 + py * 16 + pxo) = o;
  }
}
int launch_patchify(const float* px, const float* prompt_px, const float* prompt_mask, const float* /*labels*/,
                    __nv_bfloat16* A, int B, int img, cudaStream_t stream) {
  const long long total = 2LL * B * 3 * (2 * img) * (img / 8);
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(total) * 8 * 6, stream);
  BSEG_CHECK_CUDA(launch_pdl(patchify_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, stream, px, prompt_px, prompt_mask,
                             A, B, img));
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// two-stream merge after layer `merge_index`: h[:B] = (h[:B] + h[B:]) * 0.5   (modeling_seggpt.py:476-479)
// ----------------------------------------------------------------------------------------------
__global__ void merge_streams_kernel(float4* __restrict__ h, long long n4_half) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4_half;
       i += (long long)gridDim.x * blockDim.x) {
    float4 a = h[i];
    const float4 b = h[i + n4_half];
    a.x = (a.x + b.x) * 0.5f; a.y = (a.y + b.y) * 0.5f; a.z = (a.z + b.z) * 0.5f; a.w = (a.w + b.w) * 0.5f;
    h[i] = a;
  }
}
int launch_merge_streams(float* h, long long n_half, cudaStream_t stream) {
  BSEG_REQUIRE(n_half % 4 == 0, "merge_streams: size must be a multiple of 4");
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(n_half) * 12, stream);
  BSEG_CHECK_CUDA(launch_pdl(merge_streams_kernel, dim3(blocks_for(n_half / 4, 256)), dim3(256), 0, stream,
                             reinterpret_cast<float4*>(h), n_half / 4));
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// feature ensemble (modeling_seggpt.py:420-432): attention output (after proj, before the residual) of the
// query half (bottom 28 token rows) is replaced by its mean over the prompts of the same tile, then added
// to the residual stream.  nseq = nstreams * G * P sequences, sample index = (s*G + g)*P + p.
//   cross_stream == 0 : mean over p               (layers != merge_index; per stream for layers < merge_index)
//   cross_stream == 1 : mean over (s, p)          (layer == merge_index: HF averages dim 0 of the 2B batch)
// ----------------------------------------------------------------------------------------------
__global__ void ensemble_residual_kernel(float* __restrict__ h, const float* __restrict__ attn, int nstreams, int G,
                                         int P, int cross_stream, int T, int D) {
  const long long per_seq4 = (long long)T * D / 4;
  const long long half4 = per_seq4 / 2;
  const long long total = (long long)nstreams * G * per_seq4;  // one thread per (s, g, token-elem4), loops over p
  float4* h4 = reinterpret_cast<float4*>(h);
  const float4* a4 = reinterpret_cast<const float4*>(attn);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx % per_seq4;
    const long long sg = idx / per_seq4;
    const int g = static_cast<int>(sg % G);
    const int s = static_cast<int>(sg / G);
    if (e < half4) {  // prompt half: plain residual
      for (int p = 0; p < P; ++p) {
        const long long off = (((long long)s * G + g) * P + p) * per_seq4 + e;
        float4 r = h4[off];
        const float4 a = a4[off];
        r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
        h4[off] = r;
      }
    } else {
      float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
      const int s_lo = cross_stream ? 0 : s, s_hi = cross_stream ? nstreams : s + 1;
      for (int ss = s_lo; ss < s_hi; ++ss)
        for (int p = 0; p < P; ++p) {
          const float4 a = a4[(((long long)ss * G + g) * P + p) * per_seq4 + e];
          m.x += a.x; m.y += a.y; m.z += a.z; m.w += a.w;
        }
      const float inv = 1.0f / static_cast<float>((s_hi - s_lo) * P);
      m.x *= inv; m.y *= inv; m.z *= inv; m.w *= inv;
      for (int p = 0; p < P; ++p) {
        const long long off = (((long long)s * G + g) * P + p) * per_seq4 + e;
        float4 r = h4[off];
        r.x += m.x; r.y += m.y; r.z += m.z; r.w += m.w;
        h4[off] = r;
      }
    }
  }
}
int launch_ensemble_residual(float* h, const float* attn, int nstreams, int G, int P, int cross_stream, int T, int D,
                             cudaStream_t stream) {
  const long long total = (long long)nstreams * G * T * D / 4;
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(total) * P * 48, stream);
  ensemble_residual_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(h, attn, nstreams, G, P, cross_stream, T, D);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// pred_masks.mean(dim=0) over the prompts of one tile (predict_no_prompt.py:298)
__global__ void mean_over_group_kernel(const float* __restrict__ pred, float* __restrict__ out, int P, long long per,
                                       long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long g = idx / per, e = idx % per;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += pred[(g * P + p) * per + e];
    out[idx] = s / static_cast<float>(P);
  }
}
int launch_mean_over_group(const float* pred, float* out, int ngroups, int group, long long per,
                           cudaStream_t stream) {
  const long long total = ngroups * per;
  mean_over_group_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(pred, out, group, per, total);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// torch_apply_mask_rgb + normalize (src/util/ml_util.py:114-132, src/data.py:226-229,342-343):
//   out[b,c,y,x] = (palette[b, mask[b,y,x], c] / 255 - mean[c]) / std[c]
// ----------------------------------------------------------------------------------------------
__global__ void colorize_norm_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ palette,
                                     int ncls, float m0, float m1, float m2, float s0, float s1, float s2,
                                     float* __restrict__ out, int B, int HW) {
  const long long total = (long long)B * HW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = static_cast<int>(idx / HW);
    const int p = static_cast<int>(idx % HW);
    int cls = mask[idx];
    if (cls >= ncls) cls = ncls - 1;
    const uint8_t* pal = palette + ((long long)b * ncls + cls) * 3;
    float* o = out + (long long)b * 3 * HW + p;
    o[0] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(pal[0]), 255.0f), m0), s0);
    o[HW] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(pal[1]), 255.0f), m1), s1);
    o[2 * HW] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(pal[2]), 255.0f), m2), s2);
  }
}
int launch_colorize_norm(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean,
                         const float* stdv, float* out, int B, int H, int W, cudaStream_t stream) {
  const long long total = (long long)B * H * W;
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(total) * 13, stream);
  colorize_norm_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(mask, palette, num_classes, mean[0], mean[1],
                                                                   mean[2], stdv[0], stdv[1], stdv[2], out, B, H * W);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// SegGptImageProcessor.preprocess for segmentation-map prompts (HF:image_processing_seggpt.py:100-131,175-215):
// mask_to_rgb with build_palette(num_labels) -> resize NEAREST to 448 (torch interpolate index, passed as a table)
// -> (rgb - 255 mean) / (255 std).  mask uint8 [B,Hin,Win]; palette uint8 [ncls,3] shared by the batch;
// idx (nullable when Hin == Win == out) maps an output row/col to its source row/col.
// ----------------------------------------------------------------------------------------------
__global__ void colorize_resize_norm255_kernel(const uint8_t* __restrict__ mask, const uint8_t* __restrict__ palette,
                                               int ncls, float m0, float m1, float m2, float s0, float s1, float s2,
                                               const int* __restrict__ idx, float* __restrict__ out, int B, int Hin,
                                               int OS) {
  const long long total = (long long)B * OS * OS;
  const long long oplane = (long long)OS * OS;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = static_cast<int>(i % OS);
    const int oy = static_cast<int>((i / OS) % OS);
    const int b = static_cast<int>(i / oplane);
    const int sy = idx ? idx[oy] : oy, sx = idx ? idx[ox] : ox;
    int cls = mask[((long long)b * Hin + sy) * Hin + sx];
    if (cls >= ncls) cls = ncls - 1;
    const uint8_t* pal = palette + cls * 3;
    float* o = out + (long long)b * 3 * oplane + (long long)oy * OS + ox;
    o[0] = __fdiv_rn(__fsub_rn(static_cast<float>(pal[0]), m0), s0);
    o[oplane] = __fdiv_rn(__fsub_rn(static_cast<float>(pal[1]), m1), s1);
    o[2 * oplane] = __fdiv_rn(__fsub_rn(static_cast<float>(pal[2]), m2), s2);
  }
}
int launch_colorize_resize_norm255(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean255,
                                   const float* std255, const int* idx, float* out, int B, int Hin, int out_size,
                                   cudaStream_t stream) {
  const long long total = (long long)B * out_size * out_size;
  ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(total) * 13, stream);
  colorize_resize_norm255_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(
      mask, palette, num_classes, mean255[0], mean255[1], mean255[2], std255[0], std255[1], std255[2], idx, out, B, Hin,
      out_size);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// SegGptImageProcessor.post_process_semantic_segmentation (HF:image_processing_seggpt.py:254-321), used by
// src/predict_no_prompt.py:299-303: bottom half of pred_masks -> x*std + mean -> clip(255 x, 0, 255) -> nearest
// resize to the target size (torch interpolate index table) -> argmin_k sum_c (x_c - palette[k][c])^2 with the
// palette in 0..255 space -> nodata pixels to class 0.  Same op order as torch (mul, add, mul, clip; pow, sum).
// ----------------------------------------------------------------------------------------------
__global__ void postprocess_semantic_kernel(const float* __restrict__ pred, const float* __restrict__ palette255,
                                            int ncls, float m0, float m1, float m2, float s0, float s1, float s2,
                                            uint8_t* __restrict__ out_u8, long long* __restrict__ out_i64,
                                            const uint8_t* __restrict__ nodata, const int* __restrict__ idx, int B,
                                            int H, int W, int OS) {
  const long long total = (long long)B * OS * OS;
  const long long plane = 2LL * H * W;
  const float mean[3] = {m0, m1, m2}, stdv[3] = {s0, s1, s2};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = static_cast<int>(i % OS);
    const int oy = static_cast<int>((i / OS) % OS);
    const int b = static_cast<int>(i / ((long long)OS * OS));
    const int sy = idx ? idx[oy] : oy;
    const int sx = idx ? idx[ox] : ox;
    const float* p = pred + (long long)b * 3 * plane + (long long)(H + sy) * W + sx;
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float x = __fmul_rn(__fadd_rn(__fmul_rn(p[c * plane], stdv[c]), mean[c]), 255.0f);
      v[c] = fminf(fmaxf(x, 0.0f), 255.0f);
    }
    float best = 0.f;
    int bi = 0;
    for (int k = 0; k < ncls; ++k) {
      const float d0 = __fsub_rn(v[0], palette255[k * 3 + 0]);
      const float d1 = __fsub_rn(v[1], palette255[k * 3 + 1]);
      const float d2 = __fsub_rn(v[2], palette255[k * 3 + 2]);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
      if (k == 0 || d < best) { best = d; bi = k; }
    }
    if (nodata != nullptr && nodata[i]) bi = 0;
    if (out_u8) out_u8[i] = static_cast<uint8_t>(bi);
    if (out_i64) out_i64[i] = bi;
  }
}
int launch_postprocess_semantic(const float* pred, const float* palette255, int num_classes, const float* mean,
                                const float* stdv, uint8_t* out_u8, long long* out_i64, const uint8_t* nodata,
                                const int* idx, int B, int H, int W, int out_size, cudaStream_t stream) {
  const long long total = (long long)B * out_size * out_size;
  ProfScope prof(CAT_DECODE, 0, static_cast<double>(B) * H * W * 12 + static_cast<double>(total), stream);
  postprocess_semantic_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(
      pred, palette255, num_classes, mean[0], mean[1], mean[2], stdv[0], stdv[1], stdv[2], out_u8, out_i64, nodata, idx,
      B, H, W, out_size);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// PromptModel.process_pred_masks (src/model.py:155-175) + cv2.resize(INTER_NEAREST) (src/predict.py:258)
// + nodata zeroing (src/predict_no_prompt.py:303):
//   cls[b,y,x] = argmin_k sum_c (pred[b,c,H+sy,sx] - palette_norm[b,k,c])^2,  first minimum wins.
// pred is [B,3,2H,W] fp32 (bottom half decoded); idx (nullable) maps output row/col -> source row/col.
// ----------------------------------------------------------------------------------------------
__global__ void decode_palette_kernel(const float* __restrict__ pred, const float* __restrict__ palette_norm,
                                      int ncls, uint8_t* __restrict__ out_u8, long long* __restrict__ out_i64,
                                      const uint8_t* __restrict__ nodata, const int* __restrict__ idx, int B, int H,
                                      int W, int OS) {
  const long long total = (long long)B * OS * OS;
  const long long plane = 2LL * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = static_cast<int>(i % OS);
    const int oy = static_cast<int>((i / OS) % OS);
    const int b = static_cast<int>(i / ((long long)OS * OS));
    const int sy = idx ? idx[oy] : oy;
    const int sx = idx ? idx[ox] : ox;
    const float* p = pred + (long long)b * 3 * plane + (long long)(H + sy) * W + sx;
    const float v0 = p[0], v1 = p[plane], v2 = p[2 * plane];
    const float* pal = palette_norm + (long long)b * ncls * 3;
    float best = 0.f;
    int bi = 0;
    for (int k = 0; k < ncls; ++k) {
      const float d0 = __fsub_rn(v0, pal[k * 3 + 0]);
      const float d1 = __fsub_rn(v1, pal[k * 3 + 1]);
      const float d2 = __fsub_rn(v2, pal[k * 3 + 2]);
      // torch.pow(x,2) then sum over the 3 channels in order: ((d0^2 + d1^2) + d2^2), no FMA contraction
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
      if (k == 0 || d < best) { best = d; bi = k; }
    }
    if (nodata != nullptr && nodata[i]) bi = 0;
    if (out_u8) out_u8[i] = static_cast<uint8_t>(bi);
    if (out_i64) out_i64[i] = bi;
  }
}
// uint8 output, out_size a multiple of 4 and <= 2048, at most 8 classes: 4 consecutive output pixels per thread, one
// 32-bit store.  grid.y = sample, so the sample's palette sits in registers (the first version re-read it from global
// memory for every pixel: 77 load instructions per thread, LSU-bound at 0.27 of the HBM rate) and the nearest-neighbour
// index table in shared memory.
constexpr int kDecodeMaxCls = 8, kDecodeMaxOS = 2048;
__global__ void __launch_bounds__(256)
decode_palette_vec4_kernel(const float* __restrict__ pred, const float* __restrict__ palette_norm, int ncls,
                           uint8_t* __restrict__ out_u8, const uint8_t* __restrict__ nodata,
                           const int* __restrict__ idx, int H, int W, int OS) {
  __shared__ int idx_s[kDecodeMaxOS];
  for (int i = threadIdx.x; i < OS; i += blockDim.x) idx_s[i] = idx ? idx[i] : i;
  const int b = blockIdx.y;
  float pal[kDecodeMaxCls][3];
#pragma unroll
  for (int k = 0; k < kDecodeMaxCls; ++k)
#pragma unroll
    for (int c = 0; c < 3; ++c) pal[k][c] = k < ncls ? palette_norm[((long long)b * ncls + k) * 3 + c] : 0.f;
  __syncthreads();
  const int ow = OS >> 2;
  const int total = OS * ow;
  const long long plane = 2LL * H * W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ox = (i % ow) * 4;
    const int oy = i / ow;
    const float* row = pred + (long long)b * 3 * plane + (long long)(H + idx_s[oy]) * W;
    const long long o = ((long long)b * OS + oy) * OS + ox;
    uint32_t nd = 0;
    if (nodata != nullptr) nd = *reinterpret_cast<const uint32_t*>(nodata + o);
    float v[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int sx = idx_s[ox + j];
      v[j][0] = row[sx]; v[j][1] = row[plane + sx]; v[j][2] = row[2 * plane + sx];
    }
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float best = 0.f;
      uint32_t bi = 0;
#pragma unroll
      for (int k = 0; k < kDecodeMaxCls; ++k) {
        if (k < ncls) {
          const float d0 = __fsub_rn(v[j][0], pal[k][0]);
          const float d1 = __fsub_rn(v[j][1], pal[k][1]);
          const float d2 = __fsub_rn(v[j][2], pal[k][2]);
          // torch.pow(x,2) then sum over the 3 channels in order: ((d0^2 + d1^2) + d2^2), no FMA contraction
          const float d = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
          if (k == 0 || d < best) { best = d; bi = k; }
        }
      }
      if ((nd >> (8 * j)) & 0xFFu) bi = 0;
      packed |= bi << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(out_u8 + o) = packed;
  }
}
int launch_decode_palette(const float* pred, const float* palette_norm, int num_classes, uint8_t* out_u8,
                          long long* out_i64, const uint8_t* nodata, const int* idx, int B, int H, int W,
                          int out_size, cudaStream_t stream) {
  const long long total = (long long)B * out_size * out_size;
  ProfScope prof(CAT_DECODE, 0, static_cast<double>(B) * H * W * 12 + static_cast<double>(total), stream);
  if (out_u8 != nullptr && out_i64 == nullptr && out_size % 4 == 0 && reinterpret_cast<uintptr_t>(out_u8) % 4 == 0 &&
      (nodata == nullptr || reinterpret_cast<uintptr_t>(nodata) % 4 == 0) && num_classes <= kDecodeMaxCls &&
      out_size <= kDecodeMaxOS && B > 0) {
    const int per = out_size * (out_size / 4);
    // a block pays ~2 us of prologue (index table, palette): give every thread ~8 groups of four pixels
    int bx = (per + 256 * 8 - 1) / (256 * 8);
    if (bx < 1) bx = 1;
    decode_palette_vec4_kernel<<<dim3(bx, B), 256, 0, stream>>>(pred, palette_norm, num_classes, out_u8, nodata, idx, H, W,
                                                                out_size);
  }
  else
    decode_palette_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(pred, palette_norm, num_classes, out_u8, out_i64,
                                                                      nodata, idx, B, H, W, out_size);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Accumulator.update (src/predict.py:120-159; src/predict_no_prompt.py:163-186): clip the tile box to the
// scene and add a one-hot vote per pixel.  The reference counter is uint8 (H,W,4) that wraps at 256; here the
// four counters of a pixel are one u32 (identical byte layout).  A byte that wraps has its carry removed so
// the result equals per-byte mod-256 arithmetic.
// ----------------------------------------------------------------------------------------------
__global__ void vote_accumulate_kernel(uint32_t* __restrict__ counter, int Hs, int Ws, const uint8_t* __restrict__ cls,
                                       int n_tiles, int crop, const int* __restrict__ boxes, int use_atomics) {
  const long long per = (long long)crop * crop;
  const long long total = per * n_tiles;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = static_cast<int>(i / per);
    const int sy = static_cast<int>((i % per) / crop);
    const int sx = static_cast<int>(i % crop);
    const int xmin = boxes[t * 4 + 0], ymin = boxes[t * 4 + 1];
    const int dy = ymin + sy, dx = xmin + sx;
    if (dy < 0 || dy >= Hs || dx < 0 || dx >= Ws) continue;
    const uint32_t c = cls[i];
    if (c > 3) continue;
    uint32_t* w = counter + (long long)dy * Ws + dx;
    const uint32_t inc = 1u << (8 * c);
    if (use_atomics) {
      const uint32_t old = atomicAdd(w, inc);
      if (((old >> (8 * c)) & 0xFFu) == 0xFFu && c < 3) atomicSub(w, 1u << (8 * (c + 1)));
    } else {
      const uint32_t old = *w;
      const uint32_t byte = ((old >> (8 * c)) + 1u) & 0xFFu;
      *w = (old & ~(0xFFu << (8 * c))) | (byte << (8 * c));
    }
  }
}
// Tiles that do not overlap one another (no atomics needed), crop and scene width multiples of 4: 4 pixels per thread
// -- one u32 of class ids in, one uint4 of counters read-modify-written -- so the kernel moves full sectors instead of
// single bytes.  A tile whose xmin is not a multiple of 4 takes the scalar path for its pixels.
__global__ void vote_accumulate_vec4_kernel(uint32_t* __restrict__ counter, int Hs, int Ws,
                                            const uint8_t* __restrict__ cls, int n_tiles, int crop,
                                            const int* __restrict__ boxes) {
  const int cw = crop >> 2;
  const long long per = (long long)crop * cw;
  const long long total = per * n_tiles;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = static_cast<int>(i / per);
    const int sy = static_cast<int>((i % per) / cw);
    const int sx = static_cast<int>(i % cw) * 4;
    const int dy = boxes[t * 4 + 1] + sy, dx = boxes[t * 4 + 0] + sx;
    if (dy < 0 || dy >= Hs || dx + 3 < 0 || dx >= Ws) continue;
    const uint32_t c4 = *reinterpret_cast<const uint32_t*>(cls + ((long long)t * crop + sy) * crop + sx);
    uint32_t* w = counter + (long long)dy * Ws + dx;
    if (dx >= 0 && dx + 3 < Ws && (dx & 3) == 0) {
      uint4 v = *reinterpret_cast<uint4*>(w);
      uint32_t pv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = (c4 >> (8 * j)) & 0xFFu;
        if (c <= 3) {
          const uint32_t byte = ((pv[j] >> (8 * c)) + 1u) & 0xFFu;
          pv[j] = (pv[j] & ~(0xFFu << (8 * c))) | (byte << (8 * c));
        }
      }
      *reinterpret_cast<uint4*>(w) = make_uint4(pv[0], pv[1], pv[2], pv[3]);
    } else {
      for (int j = 0; j < 4; ++j) {
        const uint32_t c = (c4 >> (8 * j)) & 0xFFu;
        if (dx + j < 0 || dx + j >= Ws || c > 3) continue;
        const uint32_t old = w[j];
        const uint32_t byte = ((old >> (8 * c)) + 1u) & 0xFFu;
        w[j] = (old & ~(0xFFu << (8 * c))) | (byte << (8 * c));
      }
    }
  }
}
// Same, eight pixels per thread (crop % 8 == 0): one 8-byte load of class ids and two independent 16-byte
// read-modify-writes in flight per thread (the four-pixel kernel reached 0.56 of the HBM rate on 41 us launches).
__device__ __forceinline__ void vote_rmw4(uint4& v, uint32_t c4) {
  uint32_t pv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t c = (c4 >> (8 * j)) & 0xFFu;
    if (c <= 3) {
      const uint32_t byte = ((pv[j] >> (8 * c)) + 1u) & 0xFFu;
      pv[j] = (pv[j] & ~(0xFFu << (8 * c))) | (byte << (8 * c));
    }
  }
  v = make_uint4(pv[0], pv[1], pv[2], pv[3]);
}
__global__ void vote_accumulate_vec8_kernel(uint32_t* __restrict__ counter, int Hs, int Ws,
                                            const uint8_t* __restrict__ cls, int n_tiles, int crop,
                                            const int* __restrict__ boxes) {
  const int cw = crop >> 3;
  const long long per = (long long)crop * cw;
  const long long total = per * n_tiles;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = static_cast<int>(i / per);
    const int sy = static_cast<int>((i % per) / cw);
    const int sx = static_cast<int>(i % cw) * 8;
    const int dy = boxes[t * 4 + 1] + sy, dx = boxes[t * 4 + 0] + sx;
    if (dy < 0 || dy >= Hs || dx + 7 < 0 || dx >= Ws) continue;
    const uint2 c8 = *reinterpret_cast<const uint2*>(cls + ((long long)t * crop + sy) * crop + sx);
    uint32_t* w = counter + (long long)dy * Ws + dx;
    if (dx >= 0 && dx + 7 < Ws && (dx & 3) == 0) {
      uint4 v0 = *reinterpret_cast<uint4*>(w), v1 = *reinterpret_cast<uint4*>(w + 4);
      vote_rmw4(v0, c8.x);
      vote_rmw4(v1, c8.y);
      *reinterpret_cast<uint4*>(w) = v0;
      *reinterpret_cast<uint4*>(w + 4) = v1;
    } else {
      for (int j = 0; j < 8; ++j) {
        const uint32_t c = ((j < 4 ? c8.x : c8.y) >> (8 * (j & 3))) & 0xFFu;
        if (dx + j < 0 || dx + j >= Ws || c > 3) continue;
        const uint32_t old = w[j];
        const uint32_t byte = ((old >> (8 * c)) + 1u) & 0xFFu;
        w[j] = (old & ~(0xFFu << (8 * c))) | (byte << (8 * c));
      }
    }
  }
}
int launch_vote_accumulate(uint32_t* counter, int Hs, int Ws, const uint8_t* cls, int n_tiles, int crop,
                           const int* boxes, int use_atomics, cudaStream_t stream) {
  const long long total = (long long)crop * crop * n_tiles;
  if (total == 0) return 0;
  ProfScope prof(CAT_VOTE, 0, static_cast<double>(total) * 9, stream);
  const bool vec = !use_atomics && (crop % 4 == 0) && (Ws % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(counter) % 16 == 0) && (reinterpret_cast<uintptr_t>(cls) % 4 == 0);
  if (vec && crop % 8 == 0 && reinterpret_cast<uintptr_t>(cls) % 8 == 0)
    vote_accumulate_vec8_kernel<<<blocks_for(total / 8, 256), 256, 0, stream>>>(counter, Hs, Ws, cls, n_tiles, crop,
                                                                               boxes);
  else if (vec)
    vote_accumulate_vec4_kernel<<<blocks_for(total / 4, 256), 256, 0, stream>>>(counter, Hs, Ws, cls, n_tiles, crop,
                                                                               boxes);
  else
    vote_accumulate_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(counter, Hs, Ws, cls, n_tiles, crop, boxes,
                                                                       use_atomics);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// np.argmax(counter, axis=2) (src/predict.py:100): first maximum wins, untouched pixels -> 0
__device__ __forceinline__ uint32_t vote_argmax1(uint32_t w) {
  uint32_t best = w & 0xFFu, bi = 0;
#pragma unroll
  for (uint32_t k = 1; k < 4; ++k) {
    const uint32_t v = (w >> (8 * k)) & 0xFFu;
    if (v > best) { best = v; bi = k; }
  }
  return bi;
}
__global__ void vote_argmax_kernel(const uint32_t* __restrict__ counter, uint8_t* __restrict__ out, long long n) {
  // 4 pixels per thread: uint4 in, u32 out (the tail is handled by the first threads)
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 w = reinterpret_cast<const uint4*>(counter)[i];
    reinterpret_cast<uint32_t*>(out)[i] = vote_argmax1(w.x) | (vote_argmax1(w.y) << 8) | (vote_argmax1(w.z) << 16) |
                                          (vote_argmax1(w.w) << 24);
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = static_cast<uint8_t>(vote_argmax1(counter[i]));
}
__global__ void vote_argmax_scalar_kernel(const uint32_t* __restrict__ counter, uint8_t* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = static_cast<uint8_t>(vote_argmax1(counter[i]));
}
int launch_vote_argmax(const uint32_t* counter, uint8_t* out, long long n, cudaStream_t stream) {
  if (n == 0) return 0;
  ProfScope prof(CAT_VOTE, 0, static_cast<double>(n) * 5, stream);
  if (reinterpret_cast<uintptr_t>(counter) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0)
    vote_argmax_kernel<<<blocks_for((n + 3) / 4, 256), 256, 0, stream>>>(counter, out, n);
  else
    vote_argmax_scalar_kernel<<<blocks_for(n, 256), 256, 0, stream>>>(counter, out, n);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// Accumulator.update's image paste (src/predict.py:157): canvas[dy0:dy1, dx0:dx1] = img_crop[sy0:sy1, sx0:sx1] for a
// batch of tiles.  Overlapping tiles of one scene carry identical pixels (both are crops of the same composite), so
// the write order does not matter.
// ----------------------------------------------------------------------------------------------
__global__ void paste_tiles_kernel(uint8_t* __restrict__ canvas, int Hs, int Ws, const uint8_t* __restrict__ crops,
                                   int n_tiles, int crop, const int* __restrict__ boxes) {
  const long long per = (long long)crop * crop;
  const long long total = per * n_tiles;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int t = static_cast<int>(i / per);
    const int sy = static_cast<int>((i % per) / crop);
    const int sx = static_cast<int>(i % crop);
    const int dy = boxes[t * 4 + 1] + sy, dx = boxes[t * 4 + 0] + sx;
    if (dy < 0 || dy >= Hs || dx < 0 || dx >= Ws) continue;
    const uint8_t* src = crops + i * 3;
    uint8_t* dst = canvas + ((long long)dy * Ws + dx) * 3;
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
  }
}
int launch_paste_tiles(uint8_t* canvas, int Hs, int Ws, const uint8_t* crops, int n_tiles, int crop, const int* boxes,
                       cudaStream_t stream) {
  const long long total = (long long)crop * crop * n_tiles;
  if (total == 0) return 0;
  ProfScope prof(CAT_VOTE, 0, static_cast<double>(total) * 6, stream);
  paste_tiles_kernel<<<blocks_for(total, 256), 256, 0, stream>>>(canvas, Hs, Ws, crops, n_tiles, crop, boxes);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// overlay_prediction (src/util/img_util.py:98-116): Image.alpha_composite(RGB->RGBA base, class-colour layer).convert
// ("RGB").  Pillow's integer compositing (libImaging/AlphaComposite.c) for an opaque destination reduces to
//   coef1 = a * 128, coef2 = 255 * 128 - coef1, t = src * coef1 + dst * coef2 + (0x80 << 7),
//   out = (((t >> 8) + t) >> 8) >> 7
// and a == 0 (class without a colour, or class id outside the table) copies the base pixel.
// rgba: uint8 [n_classes][4] = (r, g, b, alpha) per class id.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint8_t pil_blend(uint32_t src, uint32_t dst, uint32_t a) {
  const uint32_t coef1 = a << 7, coef2 = (255u << 7) - coef1;
  const uint32_t t = src * coef1 + dst * coef2 + (0x80u << 7);
  return static_cast<uint8_t>((((t >> 8) + t) >> 8) >> 7);
}
__global__ void overlay_prediction_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ pred,
                                          const uint8_t* __restrict__ rgba, int n_classes, long long npix,
                                          uint8_t* __restrict__ out) {
  __shared__ uint32_t table[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    table[i] = i < n_classes ? reinterpret_cast<const uint32_t*>(rgba)[i] : 0u;
  __syncthreads();
  // 4 pixels (12 bytes) per thread: one u32 of class ids in, three u32 of RGB in and out
  const long long n4 = npix >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t c4 = reinterpret_cast<const uint32_t*>(pred)[i];
    uint32_t w[3] = {reinterpret_cast<const uint32_t*>(img)[3 * i], reinterpret_cast<const uint32_t*>(img)[3 * i + 1],
                     reinterpret_cast<const uint32_t*>(img)[3 * i + 2]};
    uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t e = table[(c4 >> (8 * j)) & 0xFFu];
      const uint32_t a = e >> 24;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int byte = 3 * j + c;
        const uint32_t d = (w[byte >> 2] >> (8 * (byte & 3))) & 0xFFu;
        const uint32_t v = a ? pil_blend((e >> (8 * c)) & 0xFFu, d, a) : d;
        o[byte >> 2] |= v << (8 * (byte & 3));
      }
    }
    reinterpret_cast<uint32_t*>(out)[3 * i] = o[0];
    reinterpret_cast<uint32_t*>(out)[3 * i + 1] = o[1];
    reinterpret_cast<uint32_t*>(out)[3 * i + 2] = o[2];
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const uint32_t e = table[pred[i]];
    const uint32_t a = e >> 24;
    for (int c = 0; c < 3; ++c) {
      const uint32_t d = img[3 * i + c];
      out[3 * i + c] = a ? pil_blend((e >> (8 * c)) & 0xFFu, d, a) : static_cast<uint8_t>(d);
    }
  }
}
__global__ void overlay_prediction_scalar_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ pred,
                                                 const uint8_t* __restrict__ rgba, int n_classes, long long npix,
                                                 uint8_t* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < npix;
       i += (long long)gridDim.x * blockDim.x) {
    const int cls = pred[i];
    const uint32_t a = cls < n_classes ? rgba[cls * 4 + 3] : 0u;
    for (int c = 0; c < 3; ++c) {
      const uint32_t d = img[3 * i + c];
      out[3 * i + c] = a ? pil_blend(rgba[cls * 4 + c], d, a) : static_cast<uint8_t>(d);
    }
  }
}
int launch_overlay_prediction(const uint8_t* img, const uint8_t* pred, const uint8_t* rgba, int n_classes,
                              long long npix, uint8_t* out, cudaStream_t stream) {
  if (npix == 0) return 0;
  ProfScope prof(CAT_VOTE, 0, static_cast<double>(npix) * 7, stream);
  const bool vec = ((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(pred) |
                     reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(rgba)) % 4 == 0);
  if (vec)
    overlay_prediction_kernel<<<blocks_for((npix + 3) / 4, 256), 256, 0, stream>>>(img, pred, rgba, n_classes, npix,
                                                                                    out);
  else
    overlay_prediction_scalar_kernel<<<blocks_for(npix, 256), 256, 0, stream>>>(img, pred, rgba, n_classes, npix, out);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ----------------------------------------------------------------------------------------------
// SegGptLoss of the reference (src/model.py:40-64): smooth-L1(beta) between pred_masks and [0 ; labels],
// masked by [0 ; yesdata], sum / keep.sum().  As written, `keep_mask.unsqueeze(1)` broadcasts to a BxB cross
// product: loss = sum_px (sum_j l_j[px]) * (sum_i keep_i[px]) / sum(keep).  per_sample=1 gives the intended
// sum_b l_b*keep_b / sum(keep); both are identical at B=1.  Forward and d(loss)/d(pred) in one pass.
// scratch (BSEG_LOSS_SCRATCH_FLOATS words): [0] = keep count (uint32), [1] = finished-block ticket (uint32),
// [2 + b] = loss numerator of block b.  Deterministic: the count is an integer sum, every block reduces its numerator
// in a fixed order and the last block to finish adds the per-block partials in block order (no float atomics), so
// the loss scalar is bit-reproducible run to run like the gradients are.
// ----------------------------------------------------------------------------------------------
constexpr int kLossMaxBlocks = 2048;
static_assert(2 + kLossMaxBlocks <= BSEG_LOSS_SCRATCH_FLOATS, "loss scratch");
__global__ void keep_count_kernel(const uint8_t* __restrict__ yes, long long n, uint32_t* __restrict__ scratch) {
  uint32_t c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    c += yes[i] ? 1u : 0u;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c != 0u) atomicAdd(&scratch[0], c);
}
__global__ void __launch_bounds__(256)
smooth_l1_kernel(const float* __restrict__ pred, const float* __restrict__ labels, const uint8_t* __restrict__ yes,
                 float beta, int per_sample, float* __restrict__ grad, uint32_t* __restrict__ scratch,
                 float* __restrict__ loss_out, int B, int HW) {
  // one thread per (pixel, channel) of the bottom half; loops over the batch
  __shared__ float warp_part[8];
  __shared__ bool is_last;
  const float denom = 3.0f * static_cast<float>(scratch[0]);  // keep.sum(): yesdata expanded to the 3 channels
  const float inv = 1.0f / denom;
  const long long total = 3LL * HW;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i / HW);
    const int p = static_cast<int>(i % HW);
    float ksum = 0.f;
    if (!per_sample)
      for (int b = 0; b < B; ++b) ksum += yes[(long long)b * HW + p] ? 1.f : 0.f;
    for (int b = 0; b < B; ++b) {
      const long long off = ((long long)b * 3 + c) * 2 * HW + HW + p;  // bottom half of [B,3,2H,W]
      const float d = pred[off] - labels[((long long)b * 3 + c) * HW + p];
      const float ad = fabsf(d);
      const float l = ad < beta ? 0.5f * d * d / beta : ad - 0.5f * beta;
      const float g = ad < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
      const float w = per_sample ? (yes[(long long)b * HW + p] ? 1.f : 0.f) : ksum;
      acc += l * w;
      if (grad) {
        grad[off] = g * w * inv;
        grad[off - HW] = 0.f;  // top half: label = 0, keep = 0
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  float* partial = reinterpret_cast<float*>(scratch + 2);
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += warp_part[w];
    partial[blockIdx.x] = s;
    __threadfence();
    is_last = atomicAdd(&scratch[1], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last && threadIdx.x < 32) {  // the last block to finish: partials in block order, one warp
    __threadfence();
    float s = 0.f;
    for (int b = threadIdx.x; b < static_cast<int>(gridDim.x); b += 32) s += __ldcg(&partial[b]);
    s = warp_sum(s);
    if (threadIdx.x == 0) loss_out[0] = s / denom;
  }
}
int launch_smooth_l1(const float* pred, const float* labels, const uint8_t* yesdata, float beta, int per_sample,
                     float* loss_out, float* grad_out, float* scratch, int B, int H, int W, cudaStream_t stream) {
  const int HW = H * W;
  ProfScope prof(CAT_LOSS, 0, static_cast<double>(B) * HW * (3 * 4 * 3 + 1), stream);
  uint32_t* sc = reinterpret_cast<uint32_t*>(scratch);
  BSEG_CHECK_CUDA(cudaMemsetAsync(sc, 0, 2 * sizeof(uint32_t), stream));
  keep_count_kernel<<<blocks_for((long long)B * HW, 1024), 256, 0, stream>>>(yesdata, (long long)B * HW, sc);
  smooth_l1_kernel<<<blocks_for(3LL * HW, 256, kLossMaxBlocks), 256, 0, stream>>>(pred, labels, yesdata, beta,
                                                                                   per_sample, grad_out, sc, loss_out,
                                                                                   B, HW);
  BSEG_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return 0;
}

}  // namespace bseg
