// Internal (C++) launcher declarations. The public C ABI lives in include/bseg.h.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm.cuh"

namespace bseg {

// gemm.cu
int launch_gemm(int mode, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, long long M, int N, int K,
                const GemmEpiParams& ep, cudaStream_t stream);
int launch_gemm_rows(int mode, const __nv_bfloat16* A, long long lda, const __nv_bfloat16* W, const GemmRows& gr, int N,
                     int K, const GemmEpiParams& ep, cudaStream_t stream);

// selects the CTA-pair GEMM kernel (cta_group::2) for N % 256 == 0; on < 0 only queries.  Returns the previous setting.
int gemm_set_cta_pairs(int on);
// 128 x 128 one-CTA tiles for launches too small to fill the SMs with 256-wide tiles (default on; < 0 queries)
int gemm_set_small_tiles(int on);
// residual + LayerNorm fusion level: 0 off, 1 (default) lin2 emits the next layer's norm1, 2 proj emits norm2 too; < 0 queries
int gemm_set_fused_ln(int level);

// attention.cu : fused softmax(q k^T * scale + decomposed rel-pos bias) v, one CTA per (seq, head, 128-query tile)
//   q, k : [nseq, heads, T, 64] bf16     vt : [nseq, heads, 64, T] bf16
//   relcat : [176, 64] bf16 = reversed rel_pos_h (111 rows, padded to 112) ; reversed rel_pos_w (55 rows, padded to 64)
//   out : [nseq, T, heads*64] bf16 (token-major, ready to be the A operand of the proj GEMM)
//   lse_out (nullable): log2-domain log-sum-exp per (seq, head, query), saved for the backward pass
int launch_attention(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt,
                     const __nv_bfloat16* relcat, __nv_bfloat16* out, float* lse_out, int nseq, int heads, int grid_h,
                     int grid_w, cudaStream_t stream);
// rows of the relcat table for a token grid: round16(2 gh - 1) + round16(2 gw - 1)  (176 for 56 x 28, 192 for 64 x 32)
int attention_relcat_rows(int grid_h, int grid_w);

// attention_bwd.cu : dq, dk, dv of the fused attention, written token-major into dqkv [nseq*T, 3*heads*64] bf16
//   q, k, v : [nseq*heads, T, 64] (qs pre-scaled like the forward's);  dO : token-major [nseq*T, heads*64]
//   lse, Dvec : [nseq*heads, T] fp32;  relcat8 [176,64] bf16;  bias_tab : scratch of nseq*heads*84*T floats (holds the
//   16-bit per-query tables the dq kernel hands to the dk/dv kernel)
int launch_attention_bwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* v, const __nv_bfloat16* dO,
                         const float* lse, const float* Dvec, const __nv_bfloat16* relcat, float* bias_tab,
                         __nv_bfloat16* dqkv, int nseq, int heads, cudaStream_t stream);

// decoder_conv.cu : conv3x3(64->64, pad 1) + LayerNorm(C=64) + erf-GELU + conv1x1(64->3), NHWC bf16 in, NCHW fp32 out
int launch_decoder_head(const __nv_bfloat16* x_nhwc, const __nv_bfloat16* w9 /*[9][64 out][64 in]*/,
                        const float* conv_b, const float* ln_w, const float* ln_b, const float* head_w /*[3][64]*/,
                        const float* head_b, float* pred /*[B,3,H,W]*/, int B, int H, int W, float eps,
                        int y_begin /*first image row computed*/, cudaStream_t stream);

int launch_decoder_head_bwd(const __nv_bfloat16* x_nhwc, const __nv_bfloat16* w9, const float* conv_b,
                            const float* ln_w, const float* ln_b, const float* head_w, const float* head_b,
                            const float* d_pred, __nv_bfloat16* d_conv, int B, int H, int W, int y0, float eps,
                            cudaStream_t stream);
int launch_decoder_conv_dgrad(const __nv_bfloat16* d_conv, const __nv_bfloat16* w9b, __nv_bfloat16* d_dec_rows, int B,
                              int H, int W, int y0, int y_first, cudaStream_t stream);

// backward.cu
int launch_transpose_bf16(const __nv_bfloat16* src, __nv_bfloat16* dst, int R, int C, int batch,
                          long long src_batch_stride, long long dst_batch_stride, long long ld_src, long long ld_dst,
                          cudaStream_t stream);
int launch_layernorm1024_bwd(const float* x, const float* dy, long long lddy, const float* gamma, const float* dh_in,
                             float* dh_out, __nv_bfloat16* dh_bf16, long long rows_per_batch, int nbatch,
                             int row_begin, int rows, float eps, cudaStream_t stream);
int launch_scale_f32_bf16(const float* src, float* dst, __nv_bfloat16* dst_bf16, float scale, long long n,
                          cudaStream_t stream);
int launch_unpatchify_prompt_grad(const float* dA, float* dprompt, int B, cudaStream_t stream);
int launch_attn_bwd_prep(const __nv_bfloat16* dO, const __nv_bfloat16* O, const __nv_bfloat16* vt, float* Dvec,
                         __nv_bfloat16* v, int nseq, int heads, int T, cudaStream_t stream);

// elementwise.cu
int launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t stream);
int launch_layernorm1024(const float* x, long long ldx, const float* gamma, const float* beta, __nv_bfloat16* out,
                         long long ldo, long long M, float eps, cudaStream_t stream);
int launch_patchify(const float* px, const float* prompt_px, const float* prompt_mask, const float* labels,
                    __nv_bfloat16* A, int B, int img, cudaStream_t stream);
int launch_merge_streams(float* h, long long n_half, cudaStream_t stream);
int launch_ensemble_residual(float* h, const float* attn, int nstreams, int G, int P, int cross_stream, int T, int D,
                             cudaStream_t stream);
int launch_mean_over_group(const float* pred, float* out, int ngroups, int group, long long per, cudaStream_t stream);
int launch_colorize_norm(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean,
                         const float* stdv, float* out, int B, int H, int W, cudaStream_t stream);
int launch_decode_palette(const float* pred, const float* palette_norm, int num_classes, uint8_t* out_u8,
                          long long* out_i64, const uint8_t* nodata, const int* idx, int B, int H, int W,
                          int out_size, cudaStream_t stream);
int launch_colorize_resize_norm255(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean255,
                                   const float* std255, const int* idx, float* out, int B, int Hin, int out_size,
                                   cudaStream_t stream);
int launch_postprocess_semantic(const float* pred, const float* palette255, int num_classes, const float* mean,
                                const float* stdv, uint8_t* out_u8, long long* out_i64, const uint8_t* nodata,
                                const int* idx, int B, int H, int W, int out_size, cudaStream_t stream);
int launch_preprocess_u8(const uint8_t* images, int chw, int n, int crop, const int* coef, const int* bounds, int ksize,
                         int prec, int band, int max_rows, const float* mean255, const float* std255, float* out_nchw,
                         cudaStream_t stream);
int launch_vote_accumulate(uint32_t* counter, int Hs, int Ws, const uint8_t* cls, int n_tiles, int crop,
                           const int* boxes, int use_atomics, cudaStream_t stream);
int launch_vote_argmax(const uint32_t* counter, uint8_t* out, long long n, cudaStream_t stream);
int launch_paste_tiles(uint8_t* canvas, int Hs, int Ws, const uint8_t* crops, int n_tiles, int crop, const int* boxes,
                       cudaStream_t stream);
int launch_overlay_prediction(const uint8_t* img, const uint8_t* pred, const uint8_t* rgba, int n_classes,
                              long long npix, uint8_t* out, cudaStream_t stream);
int launch_smooth_l1(const float* pred, const float* labels, const uint8_t* yesdata, float beta, int per_sample,
                     float* loss_out, float* grad_out, float* scratch, int B, int H, int W, cudaStream_t stream);
int launch_scene_stats(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, float* stats /*[4]*/,
                       unsigned int* scratch /*[4]*/, cudaStream_t stream);
int launch_ingest(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats, const int* boxes,
                  int n_tiles, int crop, const int* coef, const int* bounds, int ksize, int band, int max_rows,
                  const float* mean, const float* stdv, float* out_nchw, __nv_bfloat16* out_patch, long long patch_tile_stride,
                  uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream);
int launch_scene_stats_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, float* stats /*[4]*/,
                           unsigned int* scratch /*[4]*/, cudaStream_t stream);
// statistics of rows [row0, row1) only, left as order-preserving uint32 keys in scratch[4] (min key, 3 max keys) so that
// partial results of disjoint row ranges merge exactly with integer min / max; finalize turns merged keys into stats
int launch_scene_stats_rows(const void* scene, int is_f32, const uint8_t* nodata, int Hs, int Ws, int row0, int row1,
                            unsigned int* scratch, cudaStream_t stream);
int launch_scene_stats_finalize(const unsigned int* scratch, float* stats, cudaStream_t stream);
int launch_ingest_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats, const int* boxes,
                      int n_tiles, int crop, const int* coef, const int* bounds, int ksize, int band, int max_rows,
                      const float* mean, const float* stdv, float* out_nchw, __nv_bfloat16* out_patch,
                      long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream);
// native-resolution ingest (image_size == crop: no resize): composite -> u8 -> /255 -> normalise, [n,3,crop,crop]
int launch_ingest_native(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                         const int* boxes, int n_tiles, int crop, const float* mean, const float* stdv, float* out_nchw,
                         uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream);
int launch_ingest_native_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                             const int* boxes, int n_tiles, int crop, const float* mean, const float* stdv,
                             float* out_nchw, uint8_t* out_u8, uint8_t* out_nodata, cudaStream_t stream);
// augment.cu : kornia train augmentations (src/data.py:195-224) forward + gradient w.r.t. the image
int launch_train_aug_fwd(const float* image, const uint8_t* mask, const float* params, const int* order, const float* noise,
                         float noise_mean, float noise_std, const float* mean, const float* stdv, float* out_image,
                         uint8_t* out_mask, float* colour, int B, int H, int W, cudaStream_t stream);
int launch_train_aug_bwd(const float* image, const float* params, const int* order, const float* stdv,
                         const float* colour, const float* d_out, float* scratch, float* d_image, int B, int H, int W,
                         cudaStream_t stream);
// precise.cu : the fp32 accuracy mode (bseg_forward_f32).  Pointers into the handle's own fp32 copy of the weights.
struct F32Layer {
  const float *ln1_w, *ln1_b, *qkv_w, *qkv_b, *rel_pos_h, *rel_pos_w, *proj_w, *proj_b, *ln2_w, *ln2_b, *lin1_w, *lin1_b,
      *lin2_w, *lin2_b;
};
struct F32Weights {
  int num_layers = 0, merge_index = 0;
  int inter[4] = {0, 0, 0, 0};
  float eps = 1e-6f;
  const float* patch_w = nullptr;               // [1024, 768]
  const float* embed_tab[2] = {nullptr, nullptr};
  const F32Layer* layers = nullptr;             // host array
  const float *enc_ln_w = nullptr, *enc_ln_b = nullptr, *dec_embed_w = nullptr, *dec_embed_b = nullptr,
              *dec_conv_w = nullptr, *dec_conv_b = nullptr, *dec_ln_w = nullptr, *dec_ln_b = nullptr,
              *dec_head_w = nullptr, *dec_head_b = nullptr;
};
size_t f32_workspace_bytes(int B);
int forward_f32_impl(const F32Weights& w, const float* pixel_values, const float* prompt_pixel_values,
                     const float* prompt_masks, int B, int embedding_type, int P, void* workspace, float* pred_masks,
                     cudaStream_t stream);
// fp32 train step: forward that keeps its activations in `workspace`, and the backward to the prompt pixels
size_t f32_train_workspace_bytes(const F32Weights& w, int B);
int forward_f32_train_impl(const F32Weights& w, const float* pixel_values, const float* prompt_pixel_values,
                           const float* prompt_masks, int B, int embedding_type, void* workspace, float* pred_masks,
                           cudaStream_t stream);
int backward_f32_impl(const F32Weights& w, const float* d_pred_masks, int B, void* workspace, float* d_prompt,
                      cudaStream_t stream);
int launch_merge_mosaic(const float* data, const uint8_t* yesdata, int N, int C, int Hs, int Ws, float* mean,
                        uint8_t* nodata, cudaStream_t stream);

}  // namespace bseg
