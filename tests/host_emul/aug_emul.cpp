// TEST INFRASTRUCTURE ONLY.  Runs the per-pixel functions of beach_seg_b200/csrc/augment_math.cuh -- the very code the
// CUDA kernels of augment.cu call -- in plain loops on the CPU, so that `-m "not gpu"` tests can check the logic
// (dual-number Jacobians, HSV round trips, blur transpose, flips) against oracle/aug_ref.py without a GPU.
// Nothing under beach_seg_b200/ loads this; it is built by tests/test_aug_host_emul.py with g++.
#include "../../beach_seg_b200/csrc/augment_math.cuh"

using namespace bseg::aug;

extern "C" int emul_train_aug_fwd(const float* image, const uint8_t* mask, const float* params, const int32_t* order4,
                                  const float* noise, float noise_mean, float noise_std, const float* mean,
                                  const float* stdv, float* out_image, uint8_t* out_mask, float* colour, int B, int H,
                                  int W) {
  if (!valid_order(order4)) return -1000;
  const Order4 ord = {order4[0], order4[1], order4[2], order4[3]};
  const float m[3] = {mean[0], mean[1], mean[2]}, s[3] = {stdv[0], stdv[1], stdv[2]};
  const long long HW = (long long)H * W;
  for (long long b = 0; b < B; ++b)
    for (int p = 0; p < HW; ++p)
      color_fwd_px(image + b * 3 * HW, mask ? mask + b * HW : nullptr, params + b * kAugParams, ord, colour + b * 3 * HW,
                   mask ? out_mask + b * HW : nullptr, p, H, W);
  const bool quad = W % 4 == 0;  // the launcher's choice (torch allocations are 16-byte aligned)
  for (long long bc = 0; bc < 3LL * B; ++bc) {
    if (quad) {
      for (int q = 0; q < HW / 4; ++q)
        finish_fwd_quad(colour + bc * HW, params + (bc / 3) * kAugParams, noise ? noise + bc * HW : nullptr, noise_mean,
                        noise_std, m[bc % 3], s[bc % 3], out_image + bc * HW, q, H, W);
    } else {
      for (int p = 0; p < HW; ++p)
        finish_fwd_el(colour + bc * HW, params + (bc / 3) * kAugParams, noise ? noise + bc * HW : nullptr, noise_mean,
                      noise_std, m[bc % 3], s[bc % 3], out_image + bc * HW, p, H, W);
    }
  }
  return 0;
}

extern "C" int emul_train_aug_bwd(const float* image, const float* params, const int32_t* order4, const float* stdv,
                                  const float* colour, const float* d_out, float* scratch, float* d_image, int B, int H,
                                  int W) {
  if (!valid_order(order4)) return -1000;
  const Order4 ord = {order4[0], order4[1], order4[2], order4[3]};
  const float s[3] = {1.0f / stdv[0], 1.0f / stdv[1], 1.0f / stdv[2]};
  const long long HW = (long long)H * W;
  float* gd = scratch;
  float* gq = scratch + HW * 3 * B;
  const bool quad = W % 4 == 0;
  for (long long bc = 0; bc < 3LL * B; ++bc) {
    if (quad) {
      for (int q = 0; q < HW / 4; ++q)
        finish_bwd_quad(colour + bc * HW, params + (bc / 3) * kAugParams, d_out + bc * HW, s[bc % 3], gd + bc * HW,
                        gq + bc * HW, q, H, W);
    } else {
      for (int p = 0; p < HW; ++p)
        finish_bwd_el(colour + bc * HW, params + (bc / 3) * kAugParams, d_out + bc * HW, s[bc % 3], gd + bc * HW,
                      gq + bc * HW, p, H, W);
    }
  }
  for (long long b = 0; b < B; ++b)
    for (int p = 0; p < HW; ++p)
      color_bwd_px(image + b * 3 * HW, params + b * kAugParams, ord, gd + b * 3 * HW, gq + b * 3 * HW,
                   d_image + b * 3 * HW, p, H, W);
  return 0;
}
