// Per-pixel arithmetic of the train-time augmentation chain (see augment.cu), written once as host/device inline
// functions: the CUDA kernels in augment.cu are grid-stride loops around them, and tests/host_emul/aug_emul.cu runs the
// SAME functions on the CPU so the logic (dual-number Jacobians, HSV round trips, blur transpose) can be checked
// against the oracle without a GPU.  The shipped library only ever calls them from device code.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BSEG_HD __host__ __device__ __forceinline__
#else
#define BSEG_HD inline
#endif
#if defined(__CUDA_ARCH__)
// one rounding per torch op: no FMA contraction across the reference's separate tensor ops
#define F_ADD(a, b) __fadd_rn((a), (b))
#define F_SUB(a, b) __fsub_rn((a), (b))
#define F_MUL(a, b) __fmul_rn((a), (b))
#define F_DIV(a, b) __fdiv_rn((a), (b))
#define LDG(p) __ldg(p)
#define F_RCP(a) __frcp_rn(a)  // tangents only: values keep IEEE division
#else
#define F_ADD(a, b) ((a) + (b))
#define F_SUB(a, b) ((a) - (b))
#define F_MUL(a, b) ((a) * (b))
#define F_DIV(a, b) ((a) / (b))
#define LDG(p) (*(p))
#define F_RCP(a) (1.0f / (a))
#endif

namespace bseg {
namespace aug {

struct Order4 { int x, y, z, w; };

constexpr int kAugParams = 16;
// parameter row layout (floats)
enum : int {
  P_VFLIP = 0, P_HFLIP, P_BRIGHT /*additive: factor-1*/, P_CONTRAST, P_SATURATION, P_HUE_RAD, P_SHARP_ON, P_SHARP_F,
  P_ERASE_ON, P_ERASE_X, P_ERASE_Y, P_ERASE_W, P_ERASE_H, P_ERASE_VALUE, P_NOISE_ON, P_RESERVED
};

// ---- forward-mode dual numbers: value + N tangents.  N = 0 is the plain forward pass. ----
template <int N>
struct Dual {
  float v;
  float d[N > 0 ? N : 1];
};
#define DUAL_FOR for (int i_ = 0; i_ < N; ++i_)
template <int N> BSEG_HD Dual<N> cst(float v) { Dual<N> r; r.v = v; DUAL_FOR r.d[i_] = 0.f; return r; }
template <int N> BSEG_HD Dual<N> add(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = F_ADD(a.v, b.v); DUAL_FOR r.d[i_] = a.d[i_] + b.d[i_]; return r;
}
template <int N> BSEG_HD Dual<N> sub(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = F_SUB(a.v, b.v); DUAL_FOR r.d[i_] = a.d[i_] - b.d[i_]; return r;
}
template <int N> BSEG_HD Dual<N> mul(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = F_MUL(a.v, b.v); DUAL_FOR r.d[i_] = a.d[i_] * b.v + a.v * b.d[i_]; return r;
}
template <int N> BSEG_HD Dual<N> div(const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = F_DIV(a.v, b.v);
  const float inv = F_RCP(b.v);
  DUAL_FOR r.d[i_] = (a.d[i_] - r.v * b.d[i_]) * inv;
  return r;
}
template <int N> BSEG_HD Dual<N> addc(const Dual<N>& a, float c) { Dual<N> r = a; r.v = F_ADD(a.v, c); return r; }
template <int N> BSEG_HD Dual<N> subc(const Dual<N>& a, float c) { Dual<N> r = a; r.v = F_SUB(a.v, c); return r; }
template <int N> BSEG_HD Dual<N> rsubc(float c, const Dual<N>& a) {
  Dual<N> r; r.v = F_SUB(c, a.v); DUAL_FOR r.d[i_] = -a.d[i_]; return r;
}
template <int N> BSEG_HD Dual<N> mulc(const Dual<N>& a, float c) {
  Dual<N> r; r.v = F_MUL(a.v, c); DUAL_FOR r.d[i_] = a.d[i_] * c; return r;
}
template <int N> BSEG_HD Dual<N> divc(const Dual<N>& a, float c) {
  Dual<N> r; r.v = F_DIV(a.v, c); const float ic = 1.0f / c; DUAL_FOR r.d[i_] = a.d[i_] * ic; return r;
}
// torch.clamp(x, 0, 1): the gradient passes where 0 <= x <= 1
template <int N> BSEG_HD Dual<N> clamp01(const Dual<N>& a) {
  Dual<N> r; r.v = fminf(fmaxf(a.v, 0.f), 1.f);
  const bool pass = a.v >= 0.f && a.v <= 1.f;
  DUAL_FOR r.d[i_] = pass ? a.d[i_] : 0.f;
  return r;
}
template <int N> BSEG_HD Dual<N> sel(bool c, const Dual<N>& a, const Dual<N>& b) {
  Dual<N> r; r.v = c ? a.v : b.v; DUAL_FOR r.d[i_] = c ? a.d[i_] : b.d[i_]; return r;
}
// fmodf(x, m) for m > 0.  |x| < 2m covers every call of the chain (hues within a turn of [0, 2 pi), sector indices
// below 12); there the result is x or x -+ m, and that single subtraction is exact (Sterbenz: m <= |x| < 2m), so this
// equals fmodf bit for bit without libdevice's division loop.
BSEG_HD float fmod_pos(float x, float m) {
  const float ax = fabsf(x);
  if (ax < m) return x;
  if (ax < 2.f * m) return x < 0.f ? x + m : x - m;
  return fmodf(x, m);
}
// python / torch `%` for a positive modulus (sign of the result follows the divisor); derivative 1
BSEG_HD float pymodf(float x, float m) {
  float r = fmod_pos(x, m);
  if (r != 0.f && r < 0.f) r = F_ADD(r, m);
  return r;
}
template <int N> BSEG_HD Dual<N> pymod(const Dual<N>& a, float m) { Dual<N> r = a; r.v = pymodf(a.v, m); return r; }

constexpr float kTwoPi = 6.283185307179586f;

// kornia.color.rgb_to_hsv (h in radians); ties resolve to the first channel like torch.max / torch.min on the CPU
template <int N>
BSEG_HD void rgb_to_hsv(const Dual<N> (&c)[3], Dual<N>& h, Dual<N>& s, Dual<N>& v) {
  int imax = 0;
  if (c[1].v > c[0].v) imax = 1;
  if (c[2].v > (imax ? c[1].v : c[0].v)) imax = 2;
  int imin = 0;
  if (c[1].v < c[0].v) imin = 1;
  if (c[2].v < (imin ? c[1].v : c[0].v)) imin = 2;
  const Dual<N> mx = sel(imax == 0, c[0], sel(imax == 1, c[1], c[2]));
  const Dual<N> mn = sel(imin == 0, c[0], sel(imin == 1, c[1], c[2]));
  Dual<N> delta = sub(mx, mn);
  v = mx;
  s = div(delta, addc(mx, 1e-8f));
  if (delta.v == 0.f) delta = cst<N>(1.f);
  const Dual<N> rc = sub(mx, c[0]), gc = sub(mx, c[1]), bc = sub(mx, c[2]);
  const Dual<N> h1 = sub(bc, gc);
  const Dual<N> h2 = add(sub(rc, bc), mulc(delta, 2.f));
  const Dual<N> h3 = add(sub(gc, rc), mulc(delta, 4.f));
  Dual<N> hh = div(sel(imax == 0, h1, sel(imax == 1, h2, h3)), delta);
  hh = pymod(divc(hh, 6.f), 1.f);
  h = mulc(hh, kTwoPi);
}

// kornia.color.hsv_to_rgb
template <int N>
BSEG_HD void hsv_to_rgb(const Dual<N>& h, const Dual<N>& s, const Dual<N>& v, Dual<N> (&c)[3]) {
  const Dual<N> h6 = mulc(divc(h, kTwoPi), 6.f);
  const float hi = pymodf(floorf(h6.v), 6.f);
  const Dual<N> f = subc(pymod(h6, 6.f), hi);
  const Dual<N> p = mul(v, rsubc(1.f, s));
  const Dual<N> q = mul(v, rsubc(1.f, mul(f, s)));
  const Dual<N> t = mul(v, rsubc(1.f, mul(rsubc(1.f, f), s)));
  const int k = static_cast<int>(hi);
  //            k:   0  1  2  3  4  5
  // r = (v, q, p, p, t, v); g = (t, v, v, q, p, p); b = (p, p, t, v, v, q)
  c[0] = sel(k == 0 || k == 5, v, sel(k == 1, q, sel(k == 4, t, p)));
  c[1] = sel(k == 1 || k == 2, v, sel(k == 0, t, sel(k == 3, q, p)));
  c[2] = sel(k == 3 || k == 4, v, sel(k == 2, t, sel(k == 5, q, p)));
}

// ColorJiggle.apply_transform: brightness (additive), contrast (multiplicative), saturation, hue in the drawn order
template <int N>
BSEG_HD void color_chain(Dual<N> (&c)[3], const float* prm, Order4 order) {
  const float bright = LDG(prm + P_BRIGHT), contrast = LDG(prm + P_CONTRAST), sat = LDG(prm + P_SATURATION),
              hue = LDG(prm + P_HUE_RAD);
  const int ord[4] = {order.x, order.y, order.z, order.w};
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int op = ord[step];
    if (op == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = clamp01(addc(c[k], bright));
    } else if (op == 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = clamp01(mulc(c[k], contrast));
    } else {
      Dual<N> h, s, v;
      rgb_to_hsv(c, h, s, v);
      if (op == 2) {
        s = clamp01(mulc(s, sat));
      } else {
        h = addc(h, hue);
        h.v = fmod_pos(h.v, kTwoPi);  // torch.fmod: the sign follows the dividend
      }
      hsv_to_rgb(h, s, v, c);
    }
  }
}

BSEG_HD bool in_erase_box(const float* prm, int y, int x) {
  if (LDG(prm + P_ERASE_ON) == 0.f) return false;
  const int ex = static_cast<int>(LDG(prm + P_ERASE_X)), ey = static_cast<int>(LDG(prm + P_ERASE_Y));
  const int ew = static_cast<int>(LDG(prm + P_ERASE_W)), eh = static_cast<int>(LDG(prm + P_ERASE_H));
  return x >= ex && x < ex + ew && y >= ey && y < ey + eh;
}


// kornia.enhance.sharpness pieces shared by the forward and the backward finish passes
struct SharpEval {
  float T, conv, D;
  bool interior;
};
BSEG_HD SharpEval sharp_eval(const float* plane, int y, int x, int H, int W) {
  SharpEval e;
  e.T = plane[y * W + x];
  e.interior = y > 0 && y < H - 1 && x > 0 && x < W - 1;
  e.conv = 0.f;
  e.D = e.T;
  if (e.interior) {
    const float w1 = 1.0f / 13.0f, w5 = 5.0f / 13.0f;
    const float* r0 = plane + (y - 1) * W + x;
    const float* r1 = r0 + W;
    const float* r2 = r1 + W;
    float a = w1 * r0[-1];
    a = fmaf(w1, r0[0], a);
    a = fmaf(w1, r0[1], a);
    a = fmaf(w1, r1[-1], a);
    a = fmaf(w5, r1[0], a);
    a = fmaf(w1, r1[1], a);
    a = fmaf(w1, r2[-1], a);
    a = fmaf(w1, r2[0], a);
    a = fmaf(w1, r2[1], a);
    e.conv = a;
    e.D = fminf(fmaxf(a, 0.f), 1.f);
  }
  return e;
}

// All per-pixel functions below take pointers already offset to sample b (three planes of H*W for images, one for
// masks), that sample's parameter row `prm`, and the pixel index p = y*W + x inside a plane (H*W < 2^31).

// ---- forward, pass 1 (one pixel, three channels): flips + colour; the mask is flipped and zeroed in the erase box ----
BSEG_HD void color_fwd_px(const float* image, const uint8_t* mask, const float* prm, Order4 order, float* colour,
                          uint8_t* out_mask, int p, int H, int W) {
  const int HW = H * W;
  const int y = p / W, x = p - y * W;
  const int ys = LDG(prm + P_VFLIP) != 0.f ? H - 1 - y : y;
  const int xs = LDG(prm + P_HFLIP) != 0.f ? W - 1 - x : x;
  const float* src = image + ys * W + xs;
  Dual<0> c[3];
  c[0].v = LDG(src);
  c[1].v = LDG(src + HW);
  c[2].v = LDG(src + 2 * HW);
  color_chain<0>(c, prm, order);
  colour[p] = c[0].v;
  colour[p + HW] = c[1].v;
  colour[p + 2 * HW] = c[2].v;
  if (mask != nullptr) out_mask[p] = in_erase_box(prm, y, x) ? uint8_t(0) : mask[ys * W + xs];
}

// ---- forward, pass 2 (one element of channel plane ch): sharpen + erase + noise + normalise ----
BSEG_HD void finish_fwd_el(const float* colour_plane, const float* prm, const float* noise_plane, float noise_mean,
                           float noise_std, float mean, float stdv, float* out_plane, int p, int H, int W) {
  const int y = p / W, x = p - y * W;
  float v;
  if (LDG(prm + P_SHARP_ON) != 0.f) {
    const SharpEval e = sharp_eval(colour_plane, y, x, H, W);
    const float f = LDG(prm + P_SHARP_F);
    if (f == 0.f) {
      v = e.D;
    } else if (f == 1.f) {
      v = e.T;
    } else {
      v = F_ADD(e.D, F_MUL(F_SUB(e.T, e.D), f));
      if (!(f > 0.f && f < 1.f)) v = fminf(fmaxf(v, 0.f), 1.f);
    }
  } else {
    v = colour_plane[p];
  }
  if (in_erase_box(prm, y, x)) v = LDG(prm + P_ERASE_VALUE);
  if (LDG(prm + P_NOISE_ON) != 0.f) v = F_ADD(v, F_ADD(F_MUL(LDG(noise_plane + p), noise_std), noise_mean));
  out_plane[p] = F_DIV(F_SUB(v, mean), stdv);
}

// ---- backward, pass 1 (one element): d_out -> the gradient that reaches `colour` directly (gd) and the gradient that
// enters the 3x3 blur (gq; its transpose is applied by the gather in pass 2) ----
BSEG_HD void finish_bwd_el(const float* colour_plane, const float* prm, const float* d_out_plane, float inv_std,
                           float* gd_plane, float* gq_plane, int p, int H, int W) {
  const int y = p / W, x = p - y * W;
  float g = in_erase_box(prm, y, x) ? 0.f : d_out_plane[p] * inv_std;
  float direct = g, q = 0.f;
  if (LDG(prm + P_SHARP_ON) != 0.f) {
    const SharpEval e = sharp_eval(colour_plane, y, x, H, W);
    const float f = LDG(prm + P_SHARP_F);
    float dD;
    if (f == 0.f) {
      dD = g;
      direct = 0.f;
    } else if (f == 1.f) {
      dD = 0.f;
    } else {
      if (!(f > 0.f && f < 1.f)) {
        const float blend = F_ADD(e.D, F_MUL(F_SUB(e.T, e.D), f));
        if (!(blend >= 0.f && blend <= 1.f)) g = 0.f;
      }
      direct = g * f;
      dD = g - direct;
    }
    if (e.interior) q = (e.conv >= 0.f && e.conv <= 1.f) ? dD : 0.f;
    else direct += dD;  // on the border the "smoothed" image is the input itself
  }
  gd_plane[p] = direct;
  gq_plane[p] = q;
}

// the erase box of a sample, loaded once per thread
struct EraseBox {
  int x0, x1;
  bool row_in;  // the box is on and covers row y
};
BSEG_HD EraseBox load_erase_box(const float* prm, int y) {
  EraseBox e;
  e.x0 = static_cast<int>(LDG(prm + P_ERASE_X));
  e.x1 = e.x0 + static_cast<int>(LDG(prm + P_ERASE_W));
  const int ey = static_cast<int>(LDG(prm + P_ERASE_Y)), eh = static_cast<int>(LDG(prm + P_ERASE_H));
  e.row_in = LDG(prm + P_ERASE_ON) != 0.f && y >= ey && y < ey + eh;
  return e;
}

// ---- the two finish passes for FOUR consecutive pixels of a row (W % 4 == 0, 16-byte aligned planes): one 16-byte
// access per operand, the 3 x 6 neighbourhood loaded once, index arithmetic / parameter loads / row predicates shared.
// The per-pixel arithmetic (order of the nine FMAs included) is the scalar functions', so results are bit-identical. ----
struct alignas(16) F4 {
  float v[4];
};
BSEG_HD void load6(const float* row4, int x0, int W, float (&w)[6]) {
  const F4 c = *reinterpret_cast<const F4*>(row4);
  w[0] = x0 > 0 ? row4[-1] : 0.f;
  w[1] = c.v[0]; w[2] = c.v[1]; w[3] = c.v[2]; w[4] = c.v[3];
  w[5] = x0 + 4 < W ? row4[4] : 0.f;
}
BSEG_HD float conv9(const float (&a0)[6], const float (&a1)[6], const float (&a2)[6], int j) {
  const float w1 = 1.0f / 13.0f, w5 = 5.0f / 13.0f;
  float a = w1 * a0[j];
  a = fmaf(w1, a0[j + 1], a);
  a = fmaf(w1, a0[j + 2], a);
  a = fmaf(w1, a1[j], a);
  a = fmaf(w5, a1[j + 1], a);
  a = fmaf(w1, a1[j + 2], a);
  a = fmaf(w1, a2[j], a);
  a = fmaf(w1, a2[j + 1], a);
  a = fmaf(w1, a2[j + 2], a);
  return a;
}
// neighbourhood of a quad: T, interior flags and conv for its four pixels
struct QuadEval {
  float T[4], conv[4];
  bool interior[4];
};
BSEG_HD QuadEval quad_eval(const float* plane, int y, int x0, int H, int W, bool sharp) {
  QuadEval e;
  const float* row = plane + y * W + x0;
  const bool yin = sharp && y > 0 && y < H - 1;
  float a1[6];
  if (yin) {
    load6(row, x0, W, a1);
  } else {
    const F4 c = *reinterpret_cast<const F4*>(row);
    a1[1] = c.v[0]; a1[2] = c.v[1]; a1[3] = c.v[2]; a1[4] = c.v[3];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    e.T[j] = a1[j + 1];
    e.conv[j] = 0.f;
    e.interior[j] = yin && x0 + j > 0 && x0 + j < W - 1;
  }
  if (yin) {
    float a0[6], a2[6];
    load6(row - W, x0, W, a0);
    load6(row + W, x0, W, a2);
#pragma unroll
    for (int j = 0; j < 4; ++j) e.conv[j] = conv9(a0, a1, a2, j);
  }
  return e;
}

BSEG_HD void finish_fwd_quad(const float* colour_plane, const float* prm, const float* noise_plane, float noise_mean,
                             float noise_std, float mean, float stdv, float* out_plane, int quad, int H, int W) {
  const int W4 = W >> 2;
  const int y = quad / W4, x0 = (quad - y * W4) << 2;
  const bool sharp = LDG(prm + P_SHARP_ON) != 0.f;
  const float f = LDG(prm + P_SHARP_F);
  const QuadEval e = quad_eval(colour_plane, y, x0, H, W, sharp);
  F4 o;
  F4 nz;
  const bool noisy = LDG(prm + P_NOISE_ON) != 0.f;
  if (noisy) nz = *reinterpret_cast<const F4*>(noise_plane + y * W + x0);
  const float ev = LDG(prm + P_ERASE_VALUE);
  const EraseBox box = load_erase_box(prm, y);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = e.T[j];
    if (sharp) {
      const float D = e.interior[j] ? fminf(fmaxf(e.conv[j], 0.f), 1.f) : e.T[j];
      if (f == 0.f) {
        v = D;
      } else if (f != 1.f) {
        v = F_ADD(D, F_MUL(F_SUB(e.T[j], D), f));
        if (!(f > 0.f && f < 1.f)) v = fminf(fmaxf(v, 0.f), 1.f);
      }
    }
    if (box.row_in && x0 + j >= box.x0 && x0 + j < box.x1) v = ev;
    if (noisy) v = F_ADD(v, F_ADD(F_MUL(nz.v[j], noise_std), noise_mean));
    o.v[j] = F_DIV(F_SUB(v, mean), stdv);
  }
  *reinterpret_cast<F4*>(out_plane + y * W + x0) = o;
}

BSEG_HD void finish_bwd_quad(const float* colour_plane, const float* prm, const float* d_out_plane, float inv_std,
                             float* gd_plane, float* gq_plane, int quad, int H, int W) {
  const int W4 = W >> 2;
  const int y = quad / W4, x0 = (quad - y * W4) << 2;
  const bool sharp = LDG(prm + P_SHARP_ON) != 0.f;
  const float f = LDG(prm + P_SHARP_F);
  const F4 dv = *reinterpret_cast<const F4*>(d_out_plane + y * W + x0);
  QuadEval e;
  if (sharp) e = quad_eval(colour_plane, y, x0, H, W, true);
  F4 od, oq;
  const EraseBox box = load_erase_box(prm, y);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float g = (box.row_in && x0 + j >= box.x0 && x0 + j < box.x1) ? 0.f : dv.v[j] * inv_std;
    float direct = g, q = 0.f;
    if (sharp) {
      const float D = e.interior[j] ? fminf(fmaxf(e.conv[j], 0.f), 1.f) : e.T[j];
      float dD;
      if (f == 0.f) {
        dD = g;
        direct = 0.f;
      } else if (f == 1.f) {
        dD = 0.f;
      } else {
        if (!(f > 0.f && f < 1.f)) {
          const float blend = F_ADD(D, F_MUL(F_SUB(e.T[j], D), f));
          if (!(blend >= 0.f && blend <= 1.f)) g = 0.f;
        }
        direct = g * f;
        dD = g - direct;
      }
      if (e.interior[j]) q = (e.conv[j] >= 0.f && e.conv[j] <= 1.f) ? dD : 0.f;
      else direct += dD;
    }
    od.v[j] = direct;
    oq.v[j] = q;
  }
  *reinterpret_cast<F4*>(gd_plane + y * W + x0) = od;
  *reinterpret_cast<F4*>(gq_plane + y * W + x0) = oq;
}

// ---- backward, pass 2 (one pixel): blur^T gather, then J^T of the colour chain (forward-mode duals, three tangents),
// written back through the flips ----
BSEG_HD void color_bwd_px(const float* image, const float* prm, Order4 order, const float* gd, const float* gq,
                          float* d_image, int p, int H, int W) {
  const int HW = H * W;
  const int y = p / W, x = p - y * W;
  const bool sharp = LDG(prm + P_SHARP_ON) != 0.f;
  float dT[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float* gqk = gq + k * HW;
    float acc = gd[k * HW + p];
    if (sharp) {
      const float w1 = 1.0f / 13.0f, w5 = 5.0f / 13.0f;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = x + dx;
          if (xx < 0 || xx >= W) continue;
          acc = fmaf((dy == 0 && dx == 0) ? w5 : w1, gqk[yy * W + xx], acc);
        }
      }
    }
    dT[k] = acc;
  }
  const int ys = LDG(prm + P_VFLIP) != 0.f ? H - 1 - y : y;
  const int xs = LDG(prm + P_HFLIP) != 0.f ? W - 1 - x : x;
  const int soff = ys * W + xs;
  Dual<3> c[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    c[k].v = LDG(image + soff + k * HW);
    c[k].d[0] = k == 0 ? 1.f : 0.f;
    c[k].d[1] = k == 1 ? 1.f : 0.f;
    c[k].d[2] = k == 2 ? 1.f : 0.f;
  }
  color_chain<3>(c, prm, order);
#pragma unroll
  for (int j = 0; j < 3; ++j)
    d_image[soff + j * HW] = dT[0] * c[0].d[j] + dT[1] * c[1].d[j] + dT[2] * c[2].d[j];
}

inline bool valid_order(const int* o) {
  int seen = 0;
  for (int i = 0; i < 4; ++i) {
    if (o[i] < 0 || o[i] > 3) return false;
    seen |= 1 << o[i];
  }
  return seen == 15;
}

}  // namespace aug
}  // namespace bseg
