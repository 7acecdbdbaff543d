"""CPU oracle for the train-time augmentation chain of the reference.  TEST
INFRASTRUCTURE ONLY (see oracle/seggpt_ref.py for who may import this).

*** PARITY UNPINNED ***  The arithmetic lives in a third-party dependency that is neither under /root/reference nor
installed in this image: **kornia** (`environment.yml` names it without a version; no lock file).  The reference's
call site is `src/data.py:195-224`:

    K.AugmentationSequential(RandomVerticalFlip(p), RandomHorizontalFlip(p), ColorJiggle(hue, saturation, contrast,
        brightness), RandomSharpness(sharpness, p), RandomErasing(scale, p), RandomGaussianNoise(mean, std, p),
        Normalize(mean, std), data_keys=None)

applied to `{"image", "mask", ...}` dicts at `src/model.py:205` (the prompt stack, inside the autograd chain to the
prompt parameters) and `src/data.py:295-313` (the training batch).  What follows restates the published algorithms of
kornia 0.7.x (`kornia/enhance/adjust.py`, `kornia/color/hsv.py`, `kornia/augmentation/_2d/intensity/{color_jiggle,
sharpness,erasing,gaussian_noise}.py`, `kornia/augmentation/random_generator/_2d/{color_jiggle,rectangle_earse}.py`) in
plain torch, each function naming the kornia function it follows.  There are no golden vectors to pin it against
(the reference has none and kornia cannot be run here), so the parity claim for this row is "CUDA kernel == this
restatement", and the restatement itself is checked only through properties (tests/test_oracle_aug.py: identity
parameters, HSV round trip, flips are involutions, torchvision's `adjust_sharpness` == the sharpness restatement).

Random draws: kornia samples from its own `torch.distributions` objects; the order and count of draws is a kornia
implementation detail and is NOT reproduced.  Parameters are therefore explicit inputs here (`AugParams`), drawn by
`beach_seg_b200.augment.TrainAug.sample_params` with the distributions kornia documents.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


@dataclass
class AugParams:
    """One draw of every random quantity of the chain for a batch of B samples."""
    vflip: torch.Tensor          # bool [B]
    hflip: torch.Tensor          # bool [B]
    brightness: torch.Tensor     # float [B]  ColorJiggle brightness_factor (applied additively as factor - 1)
    contrast: torch.Tensor       # float [B]  contrast_factor (multiplicative)
    saturation: torch.Tensor     # float [B]  saturation_factor
    hue: torch.Tensor            # float [B]  hue_factor in turns (multiplied by 2 pi when applied)
    order: tuple                 # permutation of (0 brightness, 1 contrast, 2 saturation, 3 hue), one per call
    sharp_apply: torch.Tensor    # bool [B]
    sharp_factor: torch.Tensor   # float [B]
    erase_apply: torch.Tensor    # bool [B]
    erase_box: torch.Tensor      # int [B,4] = x, y, width, height
    erase_value: float
    noise_apply: torch.Tensor    # bool [B]
    noise: torch.Tensor          # float [B,3,H,W] standard normal
    noise_mean: float
    noise_std: float


# ---- kornia/color/hsv.py --------------------------------------------------------------------------------------
def rgb_to_hsv(image: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    max_rgb, argmax_rgb = image.max(-3)
    min_rgb, _ = image.min(-3)
    deltac = max_rgb - min_rgb
    v = max_rgb
    s = deltac / (max_rgb + eps)
    deltac = torch.where(deltac == 0, torch.ones_like(deltac), deltac)
    rc, gc, bc = torch.unbind(max_rgb.unsqueeze(-3) - image, dim=-3)
    h1 = bc - gc
    h2 = (rc - bc) + 2.0 * deltac
    h3 = (gc - rc) + 4.0 * deltac
    h = torch.stack((h1, h2, h3), dim=-3) / deltac.unsqueeze(-3)
    h = torch.gather(h, dim=-3, index=argmax_rgb.unsqueeze(-3)).squeeze(-3)
    h = (h / 6.0) % 1.0
    h = 2.0 * math.pi * h
    return torch.stack((h, s, v), dim=-3)


def hsv_to_rgb(image: torch.Tensor) -> torch.Tensor:
    h = image[..., 0, :, :] / (2 * math.pi)
    s = image[..., 1, :, :]
    v = image[..., 2, :, :]
    hi = torch.floor(h * 6) % 6
    f = ((h * 6) % 6) - hi
    one = torch.tensor(1.0)
    p = v * (one - s)
    q = v * (one - f * s)
    t = v * (one - (one - f) * s)
    hi = hi.long()
    indices = torch.stack([hi, hi + 6, hi + 12], dim=-3)
    out = torch.stack((v, q, p, p, t, v, t, v, v, q, p, p, p, p, t, v, v, q), dim=-3)
    return torch.gather(out, -3, indices)


# ---- kornia/enhance/adjust.py ---------------------------------------------------------------------------------
def _per_sample(f: torch.Tensor) -> torch.Tensor:
    return f.to(torch.float32).view(-1, 1, 1, 1)


def adjust_brightness(image, factor):
    """adjust_brightness: additive, clipped to [0,1]."""
    return (image + _per_sample(factor)).clamp(0.0, 1.0)


def adjust_contrast(image, factor):
    """adjust_contrast (the multiplicative one ColorJiggle uses; ColorJitter uses the mean-subtracting variant)."""
    return (image * _per_sample(factor)).clamp(0.0, 1.0)


def adjust_saturation(image, factor):
    hsv = rgb_to_hsv(image)
    h, s, v = torch.chunk(hsv, 3, dim=-3)
    s = torch.clamp(s * _per_sample(factor), 0.0, 1.0)
    return hsv_to_rgb(torch.cat([h, s, v], dim=-3))


def adjust_hue(image, factor_rad):
    hsv = rgb_to_hsv(image)
    h, s, v = torch.chunk(hsv, 3, dim=-3)
    h = torch.fmod(h + _per_sample(factor_rad), 2 * math.pi)
    return hsv_to_rgb(torch.cat([h, s, v], dim=-3))


def color_jiggle(image, p: AugParams):
    """ColorJiggle.apply_transform: the four adjustments in the drawn order, per-sample factors, p = 1."""
    ops = [
        lambda x: adjust_brightness(x, p.brightness - 1),
        lambda x: adjust_contrast(x, p.contrast),
        lambda x: adjust_saturation(x, p.saturation),
        lambda x: adjust_hue(x, p.hue * 2 * math.pi),
    ]
    for i in p.order:
        image = ops[int(i)](image)
    return image


def sharpness(image: torch.Tensor, factor: torch.Tensor) -> torch.Tensor:
    """kornia.enhance.sharpness: 3x3 smoothing [[1,1,1],[1,5,1],[1,1,1]]/13 on the interior (clamped to [0,1]), borders
    keep the input; blend `smooth + (input - smooth) * factor`, clamped only when factor is outside (0,1)."""
    C = image.shape[1]
    kernel = torch.tensor([[1, 1, 1], [1, 5, 1], [1, 1, 1]], dtype=image.dtype).view(1, 1, 3, 3).repeat(C, 1, 1, 1) / 13
    degenerate = torch.clamp(F.conv2d(image, kernel, bias=None, padding=0, groups=C), 0.0, 1.0)
    mask = F.pad(torch.ones_like(degenerate), [1, 1, 1, 1])
    result = torch.where(mask == 1, F.pad(degenerate, [1, 1, 1, 1]), image)
    outs = []
    for i in range(image.shape[0]):
        f = float(factor[i])
        if f == 0.0:
            outs.append(result[i])
        elif f == 1.0:
            outs.append(image[i])
        else:
            res = result[i] + (image[i] - result[i]) * factor[i].to(image.dtype)
            outs.append(res if 0.0 < f < 1.0 else torch.clamp(res, 0, 1))
    return torch.stack(outs)


def erase_boxes_mask(p: AugParams, H: int, W: int) -> torch.Tensor:
    """bbox_generator + bbox_to_mask: rows y..y+h-1, columns x..x+w-1 of the samples RandomErasing selected."""
    m = torch.zeros((p.erase_apply.shape[0], H, W), dtype=torch.bool)
    for b in range(m.shape[0]):
        if bool(p.erase_apply[b]):
            x, y, w, h = (int(v) for v in p.erase_box[b])
            m[b, max(y, 0):y + h, max(x, 0):x + w] = True
    return m


def train_aug(image: torch.Tensor, mask: torch.Tensor | None, p: AugParams, mean=IMAGE_MEAN, std=IMAGE_STD):
    """The whole chain of src/data.py:195-224 for one parameter draw.  image float32 [B,3,H,W] in [0,1]; mask
    integer [B,H,W] or None.  Geometric ops (flips) and the erasing box act on the mask too (kornia zeroes erased
    mask pixels, `RandomErasing.apply_transform_mask`); intensity ops leave it alone."""
    B, _, H, W = image.shape
    sel = lambda flag, a, b: torch.where(flag.view(-1, 1, 1, 1), a, b)
    x = sel(p.vflip, image.flip(-2), image)
    x = sel(p.hflip, x.flip(-1), x)
    x = color_jiggle(x, p)
    if bool(p.sharp_apply.any()):
        x = sel(p.sharp_apply, sharpness(x, p.sharp_factor), x)
    em = erase_boxes_mask(p, H, W)
    x = torch.where(em[:, None], torch.full_like(x, p.erase_value), x)
    x = sel(p.noise_apply, x + p.noise * p.noise_std + p.noise_mean, x)
    m_t = torch.tensor(mean, dtype=x.dtype).view(1, 3, 1, 1)
    s_t = torch.tensor(std, dtype=x.dtype).view(1, 3, 1, 1)
    x = (x - m_t) / s_t
    out_mask = None
    if mask is not None:
        m = mask
        m = torch.where(p.vflip.view(-1, 1, 1), m.flip(-2), m)
        m = torch.where(p.hflip.view(-1, 1, 1), m.flip(-1), m)
        out_mask = torch.where(em, torch.zeros_like(m), m)
    return x, out_mask
