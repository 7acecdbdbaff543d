// Train-time augmentation chain of the reference (src/data.py:195-224: kornia RandomVerticalFlip, RandomHorizontalFlip,
// ColorJiggle, RandomSharpness, RandomErasing, RandomGaussianNoise, Normalize), forward AND the gradient with respect
// to the input image: the chain sits inside the autograd path from the loss to the prompt parameters
// (src/model.py:203-207).  All random quantities are explicit inputs (one row of kAugParams floats per sample, drawn on
// the host), so the kernels are deterministic functions; the arithmetic follows kornia's published ops one rounding at
// a time (augment_math.cuh; oracle/aug_ref.py names each op).
//
// Bandwidth-bound: two passes over [B,3,H,W] fp32 each way.
//   forward : colour pass  (flip gather + 4 per-pixel colour ops)            12 B read + 12 B write (+2 B mask) per pixel
//             finish pass  (3x3 sharpen, erase box, noise, normalise)        12 (+12 noise) B read + 12 B write
//   backward: finish pass  (d_out -> direct / through-the-blur gradients)    24 B read + 24 B write
//             colour pass  (3x3 gather of the blur gradient, then J^T of the colour chain by forward-mode duals)
#include "common.cuh"
#include "host_utils.h"
#include "kernels.h"
#include "augment_math.cuh"

namespace bseg {

using namespace aug;

namespace {

struct F3 {
  float v[3];
};

__global__ void aug_color_fwd_kernel(const float* __restrict__ image, const uint8_t* __restrict__ mask,
                                     const float* __restrict__ params, Order4 order, float* __restrict__ colour,
                                     uint8_t* __restrict__ out_mask, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x)
    color_fwd_px(image, mask, params, order, colour, out_mask, idx, H, W);
}

__global__ void aug_finish_fwd_kernel(const float* __restrict__ colour, const float* __restrict__ params,
                                      const float* __restrict__ noise, float noise_mean, float noise_std, F3 mean,
                                      F3 stdv, float* __restrict__ out, int B, int H, int W) {
  const long long total = (long long)B * 3 * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x)
    finish_fwd_el(colour, params, noise, noise_mean, noise_std, mean.v, stdv.v, out, idx, H, W);
}

__global__ void aug_finish_bwd_kernel(const float* __restrict__ colour, const float* __restrict__ params,
                                      const float* __restrict__ d_out, F3 stdv, float* __restrict__ gd,
                                      float* __restrict__ gq, int B, int H, int W) {
  const long long total = (long long)B * 3 * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x)
    finish_bwd_el(colour, params, d_out, stdv.v, gd, gq, idx, H, W);
}

__global__ void aug_color_bwd_kernel(const float* __restrict__ image, const float* __restrict__ params, Order4 order,
                                     const float* __restrict__ gd, const float* __restrict__ gq,
                                     float* __restrict__ d_image, int B, int H, int W) {
  const long long total = (long long)B * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x)
    color_bwd_px(image, params, order, gd, gq, d_image, idx, H, W);
}

}  // namespace

int launch_train_aug_fwd(const float* image, const uint8_t* mask, const float* params, const int* order, const float* noise,
                         float noise_mean, float noise_std, const float* mean, const float* stdv, float* out_image,
                         uint8_t* out_mask, float* colour, int B, int H, int W, cudaStream_t stream) {
  if (!valid_order(order)) {
    set_error("train_aug: order must be a permutation of 0..3");
    return -1000;
  }
  const long long px = (long long)B * H * W;
  const Order4 ord = {order[0], order[1], order[2], order[3]};
  const F3 m = {{mean[0], mean[1], mean[2]}}, s = {{stdv[0], stdv[1], stdv[2]}};
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 26, stream);
    aug_color_fwd_kernel<<<blocks_for_px(px), 256, 0, stream>>>(image, mask, params, ord, colour, out_mask, B, H, W);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 36, stream);
    aug_finish_fwd_kernel<<<blocks_for_px(px * 3), 256, 0, stream>>>(colour, params, noise, noise_mean, noise_std, m, s,
                                                                    out_image, B, H, W);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

int launch_train_aug_bwd(const float* image, const float* params, const int* order, const float* stdv,
                         const float* colour, const float* d_out, float* scratch, float* d_image, int B, int H, int W,
                         cudaStream_t stream) {
  if (!valid_order(order)) {
    set_error("train_aug: order must be a permutation of 0..3");
    return -1000;
  }
  const long long px = (long long)B * H * W;
  const Order4 ord = {order[0], order[1], order[2], order[3]};
  const F3 s = {{stdv[0], stdv[1], stdv[2]}};
  float* gd = scratch;
  float* gq = scratch + px * 3;
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 48, stream);
    aug_finish_bwd_kernel<<<blocks_for_px(px * 3), 256, 0, stream>>>(colour, params, d_out, s, gd, gq, B, H, W);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 48, stream);
    aug_color_bwd_kernel<<<blocks_for_px(px), 256, 0, stream>>>(image, params, ord, gd, gq, d_image, B, H, W);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

}  // namespace bseg
