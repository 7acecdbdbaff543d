/* bseg.h — C ABI of libbseg.so, the B200 (sm_100a) implementation of beach_seg's segmentation hot path.
 *
 * The reference (kyle-dorman/beach_seg) is pure Python: its "plugin interface" for this path is the duck-typed
 * HuggingFace call `model(pixel_values=, prompt_pixel_values=, prompt_masks=, embedding_type=, ...) -> .pred_masks`
 * returned by `src/util/ml_util.py:7-13 load_model()` plus the tensor helpers around it.  Each entry point below
 * names the reference lines it replaces.  `HF:` = transformers/models/seggpt/ (transformers 5.5.0, the un-vendored
 * dependency that holds the arithmetic).
 *
 * Conventions
 *   - plain C, no torch types; every pointer is a caller-owned DEVICE pointer unless it says "host";
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), no host sync inside;
 *   - return 0 on success, negative on error (-cudaError_t for CUDA failures, -1000 for argument errors);
 *     bseg_last_error() returns a thread-local message;
 *   - tensors are contiguous, row-major, in the layouts the reference uses (NCHW fp32 images, (B,H,W) masks).
 */
#ifndef BSEG_H_
#define BSEG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSEG_T 1568          /* tokens per stacked 896x448 image (56 x 28 patches of 16 px); native 512-px tiles: 2048 */
#define BSEG_HIDDEN 1024
#define BSEG_HEADS 16
#define BSEG_IMG 448
#define BSEG_MAX_LAYERS 48

typedef struct bseg_handle bseg_handle;

/* fp32 DEVICE pointers to the tensors of HF SegGptForImageSegmentation.state_dict() (names in comments). */
typedef struct bseg_layer_weights {
  const float* ln1_w;  const float* ln1_b;      /* layers.i.layernorm_before.{weight,bias}        [1024] */
  const float* qkv_w;  const float* qkv_b;      /* layers.i.attention.qkv.{weight,bias}   [3072,1024],[3072] */
  const float* rel_pos_h;                       /* layers.i.attention.rel_pos_h                  [111,64] */
  const float* rel_pos_w;                       /* layers.i.attention.rel_pos_w                   [55,64] */
  const float* proj_w; const float* proj_b;     /* layers.i.attention.proj.{weight,bias}  [1024,1024],[1024] */
  const float* ln2_w;  const float* ln2_b;      /* layers.i.layernorm_after.{weight,bias}         [1024] */
  const float* lin1_w; const float* lin1_b;     /* layers.i.mlp.lin1.{weight,bias}        [4096,1024],[4096] */
  const float* lin2_w; const float* lin2_b;     /* layers.i.mlp.lin2.{weight,bias}        [1024,4096],[1024] */
} bseg_layer_weights;

typedef struct bseg_weights {
  int image_size;                 /* SegGptConfig.image_size[1]: 448 (or 0) = the reference's resized path, T = 1568;
                                   * 512 / 1024 = native-resolution mode for 512- / 1024-px tiles
                                   * (SegGptConfig(image_size=(2*tile, tile)): 64 x 32 tokens, T = 2048, rel-pos tables of
                                   * 127 / 63 rows; 128 x 64 tokens, T = 8192, 255 / 127 rows): bf16 inference only */
  int num_layers;                 /* SegGptConfig.num_hidden_layers (24) */
  int merge_index;                /* SegGptConfig.merge_index (2) */
  int intermediate_indices[4];    /* SegGptConfig.intermediate_hidden_state_indices (5,11,17,23) */
  float layer_norm_eps;           /* 1e-6 */
  const float* patch_w;           /* model.embeddings.patch_embeddings.projection.weight  [1024,3,16,16] */
  const float* patch_b;           /* ...projection.bias [1024] */
  const float* mask_token;        /* model.embeddings.mask_token            [1024] */
  const float* segment_token_input;
  const float* segment_token_prompt;
  const float* type_token_semantic;
  const float* type_token_instance;
  const float* position_embeddings; /* model.embeddings.position_embeddings [1, 14*14+1, 1024] (CLS row included) */
  const bseg_layer_weights* layers; /* host array [num_layers] */
  const float* enc_ln_w; const float* enc_ln_b;     /* model.encoder.layernorm */
  const float* dec_embed_w; const float* dec_embed_b; /* decoder.decoder_embed [16384,4096],[16384] */
  const float* dec_conv_w; const float* dec_conv_b;   /* decoder.decoder_pred.conv [64,64,3,3],[64] */
  const float* dec_ln_w; const float* dec_ln_b;       /* decoder.decoder_pred.layernorm [64] */
  const float* dec_head_w; const float* dec_head_b;   /* decoder.decoder_pred.head [3,64,1,1],[3] */
} bseg_weights;

const char* bseg_last_error(void);
int bseg_version(void);

/* Replaces load_model() (src/util/ml_util.py:7-13): packs the frozen backbone once into bf16 kernel layouts
 * (weights stay [out,in] = K-major; rel-pos tables reversed and concatenated; the additive embedding table
 * bias+segment+type+bicubic(pos) precomputed, HF:modeling_seggpt.py:145-206). `w` is a host struct. */
int bseg_create(const bseg_weights* w, bseg_handle** out, void* stream);
int bseg_destroy(bseg_handle* h);

/* Bytes of scratch a forward over `batch` model samples needs. */
size_t bseg_workspace_bytes(const bseg_handle* h, int batch);

/* Replaces SegGptForImageSegmentation.forward (HF:modeling_seggpt.py:839-959) in eval mode with the default
 * bool_masked_pos.  pixel_values / prompt_pixel_values / prompt_masks: fp32 [batch,3,448,448].
 * embedding_type: 0 = "instance", 1 = "semantic".  ensemble_prompts: 0 = feature_ensemble off; P >= 1 = on, the
 * batch is batch/P tiles of P prompts each (HF averages the whole batch == one tile).
 * pred_masks: fp32 [batch,3,896,448].  For a handle created with image_size = 512 every 448 above reads 512
 * ([batch,3,512,512] in, [batch,3,1024,512] out). */
int bseg_forward(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                 const float* prompt_masks, int batch, int embedding_type, int ensemble_prompts, void* workspace,
                 size_t workspace_bytes, float* pred_masks, void* stream);

/* bseg_forward for callers that only read the query half: pred_masks[:, :, 448:, :] is bit-identical to bseg_forward's,
 * pred_masks[:, :, :448, :] is zero.  Every consumer of pred_masks in the reference reads the bottom half only
 * (process_pred_masks src/model.py:158-160, SegGptLoss src/model.py:48-57, post_process_semantic_segmentation
 * HF:image_processing_seggpt.py:284-286), so the decoder skips the prompt half: decoder_embed runs on token rows
 * 27..55, the conv head on image rows 448..895 (-7.6 ms of a 125 ms step at batch 64). */
int bseg_forward_query_half(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                            const float* prompt_masks, int batch, int embedding_type, int ensemble_prompts,
                            void* workspace, size_t workspace_bytes, float* pred_masks, void* stream);

/* Small-batch latency: the reference calls the model with batch 1 (src/data.py:287-293, src/predict.py:234), where the
 * ~190 launches of a forward (each re-encoding its CUtensorMaps on the host) cost as much host time as device time.
 * With a limit > 0, bseg_forward / bseg_forward_query_half calls with batch <= max_batch are captured into a CUDA graph
 * the second time the same (pointers, batch, flags) come in and replayed from then on (results are bit-identical: the
 * same kernels in the same order).  Returns the previous limit; 0 (the initial value) disables graphs. */
int bseg_set_graph_batch_limit(bseg_handle* h, int max_batch);

/* ---- fp32 accuracy mode: the same forward with every operand, accumulator and activation in IEEE fp32 on the CUDA
 * cores (no tensor cores, no bf16) -- logits within 1e-4 relative of the reference's fp32 CPU forward, about 30x
 * slower than bseg_forward.  bseg_enable_fp32 copies the fp32 matrices of `w` (the struct given to bseg_create) into
 * the handle once (+1.5 GB); arguments of bseg_forward_f32 are those of bseg_forward. */
int bseg_enable_fp32(bseg_handle* h, const bseg_weights* w, void* stream);
size_t bseg_workspace_bytes_f32(const bseg_handle* h, int batch);
int bseg_forward_f32(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                     const float* prompt_masks, int batch, int embedding_type, int ensemble_prompts, void* workspace,
                     size_t workspace_bytes, float* pred_masks, void* stream);
/* The train step in the accuracy mode (autograd of HF:modeling_seggpt.py under src/model.py:233-269, what
 * bseg_forward_train / bseg_backward_to_prompt compute on the tensor cores): same arguments and the same contract
 * (d_pred_masks zero for image rows < 448; gradient w.r.t. prompt_pixel_values only), every operand, accumulator and
 * saved activation IEEE fp32, no atomics (bit-reproducible).  Workspace: bseg_train_workspace_bytes_f32 (about 1.6 GB
 * per sample at 24 layers; 0 before bseg_enable_fp32).  For checking the bf16 train step, not for training at scale. */
size_t bseg_train_workspace_bytes_f32(const bseg_handle* h, int batch);
int bseg_forward_train_f32(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                           const float* prompt_masks, int batch, int embedding_type, void* workspace,
                           size_t workspace_bytes, float* pred_masks, void* stream);
int bseg_backward_to_prompt_f32(bseg_handle* h, const float* d_pred_masks, int batch, void* workspace,
                                size_t workspace_bytes, float* d_prompt_pixel_values, void* stream);

/* ---- train step (src/model.py:233-269 + Lightning's loss.backward()): the reference differentiates the frozen HF
 * module w.r.t. prompt_pixel_values only (all backbone weights have requires_grad=False, src/util/ml_util.py:9-10).
 * bseg_train_prepare packs the transposed weight copies the dgrad GEMMs need (once, +740 MB).
 * bseg_forward_train == bseg_forward (no feature ensemble) but keeps, in `workspace`, what the backward needs;
 * bseg_backward_to_prompt consumes the same workspace:  d_pred_masks fp32 [batch,3,896,448] (must be zero in the
 * prompt half, image rows < 448, as it is for the reference's loss, src/model.py:48-57) ->
 * d_prompt_pixel_values fp32 [batch,3,448,448]. */
int bseg_train_prepare(bseg_handle* h, void* stream);
size_t bseg_train_workspace_bytes(const bseg_handle* h, int batch);
int bseg_forward_train(bseg_handle* h, const float* pixel_values, const float* prompt_pixel_values,
                       const float* prompt_masks, int batch, int embedding_type, void* workspace,
                       size_t workspace_bytes, float* pred_masks, void* stream);
int bseg_backward_to_prompt(bseg_handle* h, const float* d_pred_masks, int batch, void* workspace,
                            size_t workspace_bytes, float* d_prompt_pixel_values, void* stream);

/* tif_image 4-band branch statistics (src/util/geo_util.py:459-464): stats[0] = min over valid pixels of the
 * composite, stats[1..3] = per-channel max.  scene: uint16 [4,Hs,Ws] band-planar; nodata: uint8 [Hs,Ws].
 * scratch: 4 x uint32. */
int bseg_scene_stats(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, float* stats, uint32_t* scratch,
                     void* stream);

/* The same statistics for a scene whose rows are spread over several GPUs (src/util/geo_util.py:459-464 needs them
 * scene-global): bseg_scene_stats_rows reduces rows [row0, row1) of the device scene (uint16, or float32 when is_f32)
 * into keys[4] = order-preserving uint32 encodings of (min, max0, max1, max2); keys of disjoint row ranges merge
 * exactly with an integer min (keys[0]) / max (keys[1..3]) -- e.g. one all-reduce -- and bseg_scene_stats_finalize
 * decodes merged keys into stats[4].  Rows outside [row0, row1) are not read. */
int bseg_scene_stats_rows(const void* scene, int is_f32, const uint8_t* nodata, int Hs, int Ws, int row0, int row1,
                          uint32_t* keys, void* stream);
int bseg_scene_stats_finalize(const uint32_t* keys, float* stats, void* stream);

/* tif_image + crop_tif/padded_crop + PIL BICUBIC resize to 448 + /255 + Normalize
 * (src/util/geo_util.py:454-468,297-341; src/data.py:93-124,226-229).
 * boxes: int32 [n_tiles,4] = (xmin,ymin,xmax,ymax); coef/bounds: the PIL resampling table for crop->448
 * (int32 [448,ksize] and [448,2]), built on the host by beach_seg_b200.ingest.pil_bicubic_table().
 * mean/stdv: host float[3].  Any of the outputs may be NULL:
 *   out_nchw  fp32 [n,3,448,448] normalised;  out_patch bf16 rows of the patch-embedding operand (tile stride in
 *   elements);  out_u8 uint8 [n,crop,crop,3] composite crop;  out_nodata uint8 [n,crop,crop]. */
int bseg_ingest_u16x4(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                      const int32_t* boxes, int n_tiles, int crop, const int32_t* coef, const int32_t* bounds,
                      int ksize, const float* mean, const float* stdv, float* out_nchw, void* out_patch,
                      long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, void* stream);

/* The same chain when the model runs at the tile size (native-resolution mode, image_size == crop): get_crop skips the
 * resize (src/data.py:94), so this is tif_image + crop_tif + /255 + Normalize only -- a purely HBM-bound pass (9 B in,
 * 12 B out per pixel).  out_nchw fp32 [n,3,crop,crop] (required); out_u8 / out_nodata as above, nullable. */
int bseg_ingest_native_u16x4(const uint16_t* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                             const int32_t* boxes, int n_tiles, int crop, const float* mean, const float* stdv,
                             float* out_nchw, uint8_t* out_u8, uint8_t* out_nodata, void* stream);
int bseg_ingest_native_f32x4(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                             const int32_t* boxes, int n_tiles, int crop, const float* mean, const float* stdv,
                             float* out_nchw, uint8_t* out_u8, uint8_t* out_nodata, void* stream);

/* The same two calls for a float32 scene [4,Hs,Ws] — what merge_tifs hands to tif_image (src/util/geo_util.py:385,
 * 417-420: rasters are read with out_dtype=float32 and averaged).  Values may be negative (cubic reprojection
 * overshoot); the statistics use order-preserving keys in `scratch`. */
int bseg_scene_stats_f32(const float* scene, const uint8_t* nodata, int Hs, int Ws, float* stats, uint32_t* scratch,
                         void* stream);
int bseg_ingest_f32x4(const float* scene, const uint8_t* nodata, int Hs, int Ws, const float* stats,
                      const int32_t* boxes, int n_tiles, int crop, const int32_t* coef, const int32_t* bounds,
                      int ksize, const float* mean, const float* stdv, float* out_nchw, void* out_patch,
                      long long patch_tile_stride, uint8_t* out_u8, uint8_t* out_nodata, void* stream);

/* merge_tifs accumulation (src/util/geo_util.py:410-418): rasters already on the common grid,
 * data fp32 [n_rasters,channels,Hs,Ws], yesdata uint8 [n_rasters,Hs,Ws] (0 / nonzero) ->
 * mean fp32 [channels,Hs,Ws] = sum_n(data*yes) / sum_n(yes) (0 where the weight is 0; float32 sums in raster order,
 * bit-identical to numpy) and nodata uint8 [Hs,Ws] = !any_n(yes).  yesdata is used as the weight exactly as the
 * reference does (mask values promoted to float32), so pass 0/1 masks for a plain mean. */
int bseg_merge_mosaic(const float* data, const uint8_t* yesdata, int n_rasters, int channels, int Hs, int Ws,
                      float* mean, uint8_t* nodata, void* stream);

/* ---- src/predict_no_prompt.py path: the HF image processor (HF:image_processing_seggpt.py) on the device ----
 * SegGptImageProcessor.preprocess for images / prompt images (:134-252): uint8 RGB crops [n,crop,crop,3] (HWC,
 * layout_chw = 0) or [n,3,crop,crop] -> torchvision bicubic-antialias resize to 448 on uint8 (bit-exact restatement of
 * ATen's int16-weight kernel; coef/bounds/precision from beach_seg_b200.ops.tv_bicubic_aa_table) ->
 * (x - 255 mean)/(255 std) (mean255 / std255 host float[3], HF:image_processing_backends.py:292-331) ->
 * fp32 [n,3,448,448]. */
int bseg_preprocess_u8(const uint8_t* images, int layout_chw, int n, int crop, const int32_t* coef,
                       const int32_t* bounds, int ksize, int precision_bits, const float* mean255, const float* std255,
                       float* out_nchw, void* stream);
/* preprocess for segmentation-map prompt masks (:100-131,175-215): mask uint8 [B,in,in] -> palette colour (uint8
 * [num_classes,3] = build_palette(num_labels)) -> NEAREST resize (resize_idx int32 [out] or NULL) ->
 * (rgb - 255 mean)/(255 std) -> fp32 [B,3,out,out]. */
int bseg_colorize_resize_norm255(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean255,
                                 const float* std255, const int32_t* resize_idx, float* out, int batch, int in_size,
                                 int out_size, void* stream);
/* post_process_semantic_segmentation (:254-321) [+ nodata zeroing, src/predict_no_prompt.py:303]: pred fp32
 * [B,3,2H,W] -> bottom half, x*std+mean, clip(255x,0,255), nearest resize, argmin vs palette255 fp32 [num_classes,3]. */
int bseg_postprocess_semantic(const float* pred, const float* palette255, int num_classes, const float* mean,
                              const float* stdv, uint8_t* out_u8, int64_t* out_i64, const uint8_t* nodata,
                              const int32_t* resize_idx, int batch, int H, int W, int out_size, void* stream);

/* torch_apply_mask_rgb + normalize (src/util/ml_util.py:114-132; src/model.py:210-211,238-239).
 * mask uint8 [B,H,W]; palette uint8 [B,num_classes,3]; out fp32 [B,3,H,W]. */
int bseg_colorize_norm(const uint8_t* mask, const uint8_t* palette, int num_classes, const float* mean,
                       const float* stdv, float* out, int batch, int H, int W, void* stream);

/* PromptModel.process_pred_masks (src/model.py:155-175) [+ cv2 INTER_NEAREST resize, src/predict.py:258;
 * + nodata zeroing, src/predict_no_prompt.py:303].  pred fp32 [B,3,2H,W]; palette_norm fp32 [B,num_classes,3];
 * resize_idx: int32 [out_size] source index per output row/col, or NULL (out_size == H == W);
 * nodata: uint8 [B,out_size,out_size] or NULL; out_u8 / out_i64: [B,out_size,out_size], either may be NULL. */
int bseg_decode_palette(const float* pred, const float* palette_norm, int num_classes, uint8_t* out_u8,
                        int64_t* out_i64, const uint8_t* nodata, const int32_t* resize_idx, int batch, int H, int W,
                        int out_size, void* stream);

/* pred_masks.mean(dim=0) over the prompts of a tile (src/predict_no_prompt.py:298). */
int bseg_mean_over_prompts(const float* pred, float* out, int n_tiles, int prompts, long long elems_per_sample,
                           void* stream);

/* The datamodule's `train_aug` (src/data.py:195-224: kornia RandomVerticalFlip, RandomHorizontalFlip, ColorJiggle,
 * RandomSharpness, RandomErasing, RandomGaussianNoise, Normalize), applied to the prompt stack inside the autograd
 * chain (src/model.py:203-207) and to the training batch (src/data.py:295-313).  SURVEY section 8(f) rank 2.
 * Every random quantity is an input: params fp32 [B,16] per sample =
 *   {vflip, hflip, brightness_factor-1, contrast_factor, saturation_factor, hue (radians), sharp_on, sharp_factor,
 *    erase_on, erase_x, erase_y, erase_w, erase_h, erase_value, noise_on, 0};
 * order4: HOST int32[4], the permutation of (0 brightness, 1 contrast, 2 saturation, 3 hue) ColorJiggle drew;
 * noise: fp32 [B,3,H,W] standard normal (may be NULL when no sample has noise_on).
 * image fp32 [B,3,H,W] in [0,1]; mask uint8 [B,H,W] or NULL (flipped, zeroed inside the erase box);
 * out_image fp32 [B,3,H,W]; out_mask uint8 [B,H,W] or NULL; colour_out fp32 [B,3,H,W]: the image after the colour
 * ops, which bseg_train_aug_bwd needs again (caller-owned, keep it until the backward ran). */
int bseg_train_aug_fwd(const float* image, const uint8_t* mask, const float* params, const int32_t* order4,
                       const float* noise, float noise_mean, float noise_std, const float* mean, const float* stdv,
                       float* out_image, uint8_t* out_mask, float* colour_out, int batch, int H, int W, void* stream);
/* d(out_image)/d(image)^T applied to d_out: torch autograd through the kornia chain in the reference.
 * scratch: fp32 [2,B,3,H,W]; d_image fp32 [B,3,H,W] (every element written). */
int bseg_train_aug_bwd(const float* image, const float* params, const int32_t* order4, const float* stdv,
                       const float* colour_out, const float* d_out, float* scratch, float* d_image, int batch, int H,
                       int W, void* stream);

/* Accumulator.update (src/predict.py:120-159; src/predict_no_prompt.py:163-186).  counter: uint32 [Hs,Ws]
 * == uint8 [Hs,Ws,4] votes;  cls: uint8 [n_tiles,crop,crop];  boxes: int32 [n_tiles,4].
 * use_atomics must be 1 when tiles of one call may overlap. */
int bseg_vote_accumulate(uint32_t* counter, int Hs, int Ws, const uint8_t* cls, int n_tiles, int crop,
                         const int32_t* boxes, int use_atomics, void* stream);

/* np.argmax(counter, axis=2) (src/predict.py:100; src/predict_no_prompt.py:141). out: uint8 [Hs,Ws]. */
int bseg_vote_argmax(const uint32_t* counter, uint8_t* out, long long n_pixels, void* stream);

/* Accumulator.update's image paste (src/predict.py:157): canvas uint8 [Hs,Ws,3]; crops uint8 [n_tiles,crop,crop,3];
 * boxes int32 [n_tiles,4]; the part of each crop inside the scene overwrites the canvas. */
int bseg_paste_tiles_u8(uint8_t* canvas, int Hs, int Ws, const uint8_t* crops, int n_tiles, int crop,
                        const int32_t* boxes, void* stream);

/* overlay_prediction (src/util/img_util.py:98-116): Pillow's Image.alpha_composite of the class-colour layer over the
 * RGB image, bit-exact.  img, out: uint8 [n_pixels,3]; pred: uint8 [n_pixels] class ids;
 * class_rgba: uint8 [n_classes,4] = (r,g,b,alpha), alpha 0 = class without a colour (CLASS_COLORS[c] is None). */
int bseg_overlay_prediction(const uint8_t* img, const uint8_t* pred, const uint8_t* class_rgba, int n_classes,
                            long long n_pixels, uint8_t* out, void* stream);

/* SegGptLoss (src/model.py:40-64), forward and gradient w.r.t. pred in one pass.
 * pred fp32 [B,3,2H,W]; labels fp32 [B,3,H,W]; yesdata uint8 [B,H,W]; per_sample: 0 = as written in the
 * reference (BxB broadcast), 1 = per-sample masking; loss_out: fp32 [1]; grad_out: fp32 [B,3,2H,W] or NULL;
 * scratch: BSEG_LOSS_SCRATCH_FLOATS 32-bit words (keep count, block ticket, per-block partial sums: the reduction is
 * two-stage in a fixed order with no float atomics, so loss_out is bit-reproducible run to run). */
#define BSEG_LOSS_SCRATCH_FLOATS 2052
int bseg_loss_smoothl1_fwd_bwd(const float* pred, const float* labels, const uint8_t* yesdata, float beta,
                               int per_sample, float* loss_out, float* grad_out, float* scratch, int batch, int H,
                               int W, void* stream);

/* ---- building blocks, exported for unit tests and the benchmark's roofline legs ---- */

/* GEMM kernel variant: 1 = CTA pairs (clusters of two CTAs, tcgen05.mma.cta_group::2, 256 x 256 output tiles, W tiles
 * fetched once per pair) for every GEMM with N % 256 == 0; 0 = one CTA per 128 x 256 tile.  Both accumulate in the same
 * order, so results are bit-identical.  on < 0 only queries.  Returns the previous setting (initial value: environment
 * BSEG_GEMM_2CTA, else the library default). */
int bseg_gemm_set_cta_pairs(int on);
/* Tile shape for small launches (one or a few tiles per call, the reference's own call pattern src/predict.py:234):
 * 1 (default) lets a GEMM whose 256-wide tiles would leave most SMs idle run on 128 x 128 one-CTA tiles instead;
 * 0 always uses the 256-wide tiles.  Results are bit-identical.  Returns the previous setting; < 0 only queries. */
int bseg_gemm_set_small_tiles(int on);
/* Residual + LayerNorm fusion level.  1 (default): the MLP's lin2 GEMM of every encoder layer (HF:modeling_seggpt.py:356,
 * 433-441) also writes the LayerNorm that reads its rows next (norm1 of the next layer, :420), so the fp32 residual
 * stream is not read back from HBM by a separate LayerNorm launch; 2: the attention projection (:225,420-432) also
 * writes norm2 of the layer (:433) (slower than the separate launch on B200: that GEMM's epilogue is its bottleneck);
 * 0: bseg_layernorm1024's kernel after every residual GEMM as in round 1.  The row statistics are merged in a different
 * (fixed) order, so the bf16 LayerNorm output can differ from the unfused path in the last bit.  Returns the previous
 * level; < 0 only queries. */
int bseg_gemm_set_fused_ln(int level);
/* Programmatic dependent launch for small forwards (default 1): in a bseg_forward* call of at most 2 tiles the GEMM,
 * attention, LayerNorm, patchify, merge and decoder-head kernels are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization and call griddepcontrol.wait after their prologue (barrier / TMEM /
 * tensor-map set-up), so a kernel's launch latency and prologue run under the tail of its predecessor.  Results are
 * bit-identical; a batch-1 forward (181 launches of 10-40 us, the reference's own call pattern src/predict.py:234) takes
 * 2.96 instead of 3.33 ms on B200.  Returns the previous setting; < 0 only queries. */
int bseg_set_pdl(int on);
/* One residual GEMM with the fused LayerNorm, for tests and probes:
 *   h[M,1024] (fp32, in place) += A[M,K] * W[1024,K]^T + bias ;  ln_out[M,1024] (bf16) = LayerNorm(h) * gamma + beta.
 * scratch: bseg_gemm_resid_ln_scratch_bytes(M) bytes of device memory (tagged row statistics; initialised by
 * the call). */
size_t bseg_gemm_resid_ln_scratch_bytes(long long M);
int bseg_gemm_bf16_resid_ln(const void* A, long long lda, const void* W, long long M, int K, const float* bias,
                            float* h, const float* gamma, const float* beta, void* ln_out, float eps, void* scratch,
                            void* stream);

/* D = A[M,K] * W[N,K]^T (+bias); A, W bf16; out fp32 (out_is_bf16 == 0) or bf16; gelu applies to bf16 output. */
int bseg_gemm_bf16(const void* A, long long lda, const void* W, long long M, int N, int K, const float* bias,
                   void* out, long long ldc, int out_is_bf16, int gelu, void* stream);
/* LayerNorm over 1024 columns: x fp32 [M,ldx] -> out bf16 [M,ldo]. */
int bseg_layernorm1024(const float* x, long long ldx, const float* gamma, const float* beta, void* out,
                       long long ldo, long long M, float eps, void* stream);
/* Fused rel-pos attention (HF:modeling_seggpt.py:268-348).  q bf16 [nseq,16,1568,64] = the query projection TIMES
 * head_dim^-0.5 * log2(e) (what bseg_forward's QKV epilogue writes: the score then lands in the log2 domain with no
 * per-element scaling); k bf16 [nseq,16,1568,64]; vt bf16 [nseq,16,64,1568]; relcat bf16 [176,64] from
 * bseg_pack_relcat; out bf16 [nseq,1568,1024]. */
int bseg_attention(const void* q, const void* k, const void* vt, const void* relcat, void* out, int nseq,
                   void* stream);
/* The same for a token grid grid_h x grid_w (56 x 28; 64 x 32 / 128 x 64 = native 512- / 1024-px tiles), with
 * relcat from bseg_pack_relcat_grid (bseg_relcat_rows(grid_h, grid_w) x 64 bf16); lse may be NULL. */
int bseg_attention_grid(const void* q, const void* k, const void* vt, const void* relcat, void* out, float* lse,
                        int nseq, int grid_h, int grid_w, void* stream);
int bseg_relcat_rows(int grid_h, int grid_w);
int bseg_pack_relcat_grid(const float* rel_pos_h, const float* rel_pos_w, void* relcat, int grid_h, int grid_w,
                          void* stream);
/* The same with the log2-domain log-sum-exp per (seq, head, query) written to lse fp32 [nseq,16,1568]. */
int bseg_attention_fwd_lse(const void* q, const void* k, const void* vt, const void* relcat, void* out, float* lse,
                           int nseq, void* stream);
/* Backward of the fused attention (torch autograd through HF:modeling_seggpt.py:268-348 in the reference); q as for
 * bseg_attention (pre-scaled), the gradients are those w.r.t. the UNSCALED q / k / v projections.
 * out / d_out: bf16 token-major [nseq*1568, 1024]; dqkv: bf16 [nseq*1568, 3072] = (dq | dk | dv), head-major inside. */
size_t bseg_attention_bwd_scratch_bytes(int nseq);
int bseg_attention_bwd(const void* q, const void* k, const void* vt, const void* out, const void* d_out,
                       const float* lse, const void* relcat, void* dqkv, int nseq, void* scratch, size_t scratch_bytes,
                       void* stream);
/* LayerNorm(1024) backward: dh_out = (dh_in ? dh_in : 0) + dLN(x, gamma; dy); dh_bf16 (nullable) = bf16(dh_out). */
int bseg_layernorm1024_bwd(const float* x, const float* dy, long long lddy, const float* gamma, const float* dh_in,
                           float* dh_out, void* dh_bf16, long long M, float eps, void* stream);
/* out_bf16[M,N] = (A[M,K] * W[N,K]^T) * gelu'(z[M,N])  (backward of lin1's erf-GELU; z = saved pre-activation). */
int bseg_gemm_bf16_dgelu(const void* A, long long lda, const void* W, long long M, int N, int K, const void* z,
                         void* out, long long ldc, void* stream);
/* Decoder-head backward for image rows >= y0: d_pred fp32 [B,3,H,W] -> d_conv bf16 [B,H-y0,W,64] (scratch) ->
 * d_dec_rows bf16 [B*(H/16)*(W/16), 16384] (rows of the decoder_embed dgrad operand; written for image rows
 * >= y0-16).  w9b = bseg_pack_conv_w9_dgrad(w9). */
int bseg_pack_conv_w9_dgrad(const void* w9, void* w9b, void* stream);
int bseg_decoder_head_bwd(const void* x_nhwc, const void* w9, const void* w9b, const float* conv_b, const float* ln_w,
                          const float* ln_b, const float* head_w, const float* head_b, const float* d_pred,
                          void* d_conv, void* d_dec_rows, int batch, int H, int W, int y0, float eps, void* stream);
/* relcat = 8 * [reverse(rel_pos_h) (111 rows) ; 0 ; reverse(rel_pos_w) (55 rows) ; 0...] as bf16 [176,64]
 * (8 = 1 / head_dim^-0.5: the pre-scaled q times this table is the rel-pos bias in the log2 domain). */
int bseg_pack_relcat(const float* rel_pos_h, const float* rel_pos_w, void* relcat, void* stream);
/* Decoder head: conv3x3 + LN(C) + GELU + conv1x1. x bf16 NHWC [B,H,W,64]; w9 bf16 [9,64,64]; pred fp32 [B,3,H,W]. */
int bseg_decoder_head(const void* x_nhwc, const void* w9, const float* conv_b, const float* ln_w, const float* ln_b,
                      const float* head_w, const float* head_b, float* pred, int batch, int H, int W, float eps,
                      void* stream);
int bseg_pack_conv_w9(const float* conv_w, void* w9, void* stream);
int bseg_f32_to_bf16(const float* src, void* dst, long long n, void* stream);

/* Per-launch CUDA-event timing on the launching stream, summed per kernel category (bench.py's roofline leg).
 * Categories: 0 gemm, 1 attention, 2 layernorm, 3 decoder_head, 4 ingest, 5 decode, 6 vote, 7 elementwise, 8 loss.
 * collect() synchronises the recorded events and fills HOST arrays of BSEG_PROFILE_CATEGORIES entries:
 * milliseconds, launches, algorithmic work (FLOP) and algorithmic bytes. */
#define BSEG_PROFILE_CATEGORIES 9
int bseg_profile_enable(int on);
int bseg_profile_collect(double* ms, long long* launches, double* work, double* bytes);
/* GEMM split of the last collect(): 16 entries indexed by epilogue mode (+8 when K > 2048). */
int bseg_profile_collect_gemm(double* ms, double* work);

/* number of kernel launches issued by this library since process start (bench.py's gpu_launches) */
long long bseg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BSEG_H_ */
