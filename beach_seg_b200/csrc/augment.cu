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

// Grids: x covers the H*W pixels of one plane (one element per thread, 32-bit index arithmetic), y = sample (pixel
// kernels) or sample*3 + channel (element kernels), so the parameter row is uniform per block.
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
aug_color_fwd_kernel(const float* __restrict__ image, const uint8_t* __restrict__ mask, const float* __restrict__ params,
                     Order4 order, float* __restrict__ colour, uint8_t* __restrict__ out_mask, int H, int W) {
  const int HW = H * W;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= HW) return;
  const long long b = blockIdx.y;
  color_fwd_px(image + b * 3 * HW, mask ? mask + b * HW : nullptr, params + b * kAugParams, order, colour + b * 3 * HW,
               mask ? out_mask + b * HW : nullptr, p, H, W);
}

__global__ void __launch_bounds__(kThreads)
aug_finish_fwd_kernel(const float* __restrict__ colour, const float* __restrict__ params, const float* __restrict__ noise,
                      float noise_mean, float noise_std, F3 mean, F3 stdv, float* __restrict__ out, int H, int W) {
  const int HW = H * W;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= HW) return;
  const long long bc = blockIdx.y;
  const int b = blockIdx.y / 3, ch = blockIdx.y - b * 3;
  finish_fwd_el(colour + bc * HW, params + (long long)b * kAugParams, noise ? noise + bc * HW : nullptr, noise_mean,
                noise_std, mean.v[ch], stdv.v[ch], out + bc * HW, p, H, W);
}

__global__ void __launch_bounds__(kThreads)
aug_finish_bwd_kernel(const float* __restrict__ colour, const float* __restrict__ params, const float* __restrict__ d_out,
                      F3 inv_std, float* __restrict__ gd, float* __restrict__ gq, int H, int W) {
  const int HW = H * W;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= HW) return;
  const long long bc = blockIdx.y;
  const int b = blockIdx.y / 3, ch = blockIdx.y - b * 3;
  finish_bwd_el(colour + bc * HW, params + (long long)b * kAugParams, d_out + bc * HW, inv_std.v[ch], gd + bc * HW,
                gq + bc * HW, p, H, W);
}

__global__ void __launch_bounds__(kThreads)
aug_finish_fwd_quad_kernel(const float* __restrict__ colour, const float* __restrict__ params,
                           const float* __restrict__ noise, float noise_mean, float noise_std, F3 mean, F3 stdv,
                           float* __restrict__ out, int H, int W) {
  const int HW = H * W;
  const int quad = blockIdx.x * kThreads + threadIdx.x;
  if (quad >= (HW >> 2)) return;
  const long long bc = blockIdx.y;
  const int b = blockIdx.y / 3, ch = blockIdx.y - b * 3;
  finish_fwd_quad(colour + bc * HW, params + (long long)b * kAugParams, noise ? noise + bc * HW : nullptr, noise_mean,
                  noise_std, mean.v[ch], stdv.v[ch], out + bc * HW, quad, H, W);
}

__global__ void __launch_bounds__(kThreads)
aug_finish_bwd_quad_kernel(const float* __restrict__ colour, const float* __restrict__ params,
                           const float* __restrict__ d_out, F3 inv_std, float* __restrict__ gd, float* __restrict__ gq,
                           int H, int W) {
  const int HW = H * W;
  const int quad = blockIdx.x * kThreads + threadIdx.x;
  if (quad >= (HW >> 2)) return;
  const long long bc = blockIdx.y;
  const int b = blockIdx.y / 3, ch = blockIdx.y - b * 3;
  finish_bwd_quad(colour + bc * HW, params + (long long)b * kAugParams, d_out + bc * HW, inv_std.v[ch], gd + bc * HW,
                  gq + bc * HW, quad, H, W);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(kThreads)
aug_color_bwd_kernel(const float* __restrict__ image, const float* __restrict__ params, Order4 order,
                     const float* __restrict__ gd, const float* __restrict__ gq, float* __restrict__ d_image, int H,
                     int W) {
  const int HW = H * W;
  const int p = blockIdx.x * kThreads + threadIdx.x;
  if (p >= HW) return;
  const long long b = blockIdx.y;
  color_bwd_px(image + b * 3 * HW, params + b * kAugParams, order, gd + b * 3 * HW, gq + b * 3 * HW,
               d_image + b * 3 * HW, p, H, W);
}

int check_shape(const int* order, int B, int H, int W) {
  if (!valid_order(order)) {
    set_error("train_aug: order must be a permutation of 0..3");
    return -1000;
  }
  if ((long long)H * W >= (1LL << 30) || (long long)B * 3 > 65535) {
    set_error("train_aug: plane of %d x %d pixels or batch %d too large for one launch", H, W, B);
    return -1000;
  }
  return 0;
}

}  // namespace

int launch_train_aug_fwd(const float* image, const uint8_t* mask, const float* params, const int* order, const float* noise,
                         float noise_mean, float noise_std, const float* mean, const float* stdv, float* out_image,
                         uint8_t* out_mask, float* colour, int B, int H, int W, cudaStream_t stream) {
  if (int rc = check_shape(order, B, H, W)) return rc;
  const long long px = (long long)B * H * W;
  const unsigned gx = static_cast<unsigned>((H * W + kThreads - 1) / kThreads);
  const Order4 ord = {order[0], order[1], order[2], order[3]};
  const F3 m = {{mean[0], mean[1], mean[2]}}, s = {{stdv[0], stdv[1], stdv[2]}};
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 26, stream);
    aug_color_fwd_kernel<<<dim3(gx, B), kThreads, 0, stream>>>(image, mask, params, ord, colour, out_mask, H, W);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 36, stream);
    if (W % 4 == 0 && aligned16(colour) && aligned16(out_image) && aligned16(noise)) {
      const unsigned gq4 = static_cast<unsigned>((H * W / 4 + kThreads - 1) / kThreads);
      aug_finish_fwd_quad_kernel<<<dim3(gq4, B * 3), kThreads, 0, stream>>>(colour, params, noise, noise_mean, noise_std,
                                                                          m, s, out_image, H, W);
    } else {
      aug_finish_fwd_kernel<<<dim3(gx, B * 3), kThreads, 0, stream>>>(colour, params, noise, noise_mean, noise_std, m, s,
                                                                     out_image, H, W);
    }
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

int launch_train_aug_bwd(const float* image, const float* params, const int* order, const float* stdv,
                         const float* colour, const float* d_out, float* scratch, float* d_image, int B, int H, int W,
                         cudaStream_t stream) {
  if (int rc = check_shape(order, B, H, W)) return rc;
  const long long px = (long long)B * H * W;
  const unsigned gx = static_cast<unsigned>((H * W + kThreads - 1) / kThreads);
  const Order4 ord = {order[0], order[1], order[2], order[3]};
  const F3 s = {{1.0f / stdv[0], 1.0f / stdv[1], 1.0f / stdv[2]}};
  float* gd = scratch;
  float* gq = scratch + px * 3;
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 48, stream);
    if (W % 4 == 0 && aligned16(colour) && aligned16(d_out) && aligned16(gd) && aligned16(gq)) {
      const unsigned gq4 = static_cast<unsigned>((H * W / 4 + kThreads - 1) / kThreads);
      aug_finish_bwd_quad_kernel<<<dim3(gq4, B * 3), kThreads, 0, stream>>>(colour, params, d_out, s, gd, gq, H, W);
    } else {
      aug_finish_bwd_kernel<<<dim3(gx, B * 3), kThreads, 0, stream>>>(colour, params, d_out, s, gd, gq, H, W);
    }
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  {
    ProfScope prof(CAT_ELEMENTWISE, 0, static_cast<double>(px) * 48, stream);
    aug_color_bwd_kernel<<<dim3(gx, B), kThreads, 0, stream>>>(image, params, ord, gd, gq, d_image, H, W);
    BSEG_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

}  // namespace bseg
