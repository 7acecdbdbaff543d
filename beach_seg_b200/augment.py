"""Mirror of the reference datamodule's augmentation pipelines (src/data.py:195-234), SURVEY section 8(f) rank 2.

`TrainAug(conf)` stands in for `K.AugmentationSequential(RandomVerticalFlip, RandomHorizontalFlip, ColorJiggle,
RandomSharpness, RandomErasing, RandomGaussianNoise, Normalize, data_keys=None)` and is called the same way: with a
dict holding "image" ([B,3,H,W] float in [0,1]) and "mask" ([B,H,W] or [B,1,H,W] integer); other keys pass through.
It runs on the GPU (`bseg_train_aug_fwd`) and is differentiable with respect to the image (`bseg_train_aug_bwd`),
because the reference applies it to the stack of trainable prompt images (src/model.py:203-207).  `InferenceAug` is
the `aug` pipeline (CenterCrop = identity at inpt_size, Normalize).

Random parameters are drawn on the host with the distributions kornia documents (Bernoulli(p) per sample for the
flips / sharpness / erasing / noise, U(1-b, 1+b) colour factors, one permutation of the four colour ops per call,
U(0, sharpness), RandomErasing's area / aspect-ratio / position draw); kornia's own draw ORDER is an implementation
detail of a library that is not in this image, so a seed does not reproduce kornia's stream (oracle/aug_ref.py).
There is no CPU fallback: tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch

from . import _lib
from .ops import IMAGE_MEAN, IMAGE_STD

N_PARAMS = 16
(P_VFLIP, P_HFLIP, P_BRIGHT, P_CONTRAST, P_SATURATION, P_HUE_RAD, P_SHARP_ON, P_SHARP_F, P_ERASE_ON, P_ERASE_X,
 P_ERASE_Y, P_ERASE_W, P_ERASE_H, P_ERASE_VALUE, P_NOISE_ON) = range(15)


def _order_arr(order):
    return (C.c_int32 * 4)(*[int(v) for v in order])


def _raw_fwd(fn, image, mask, params, order, noise, noise_mean, noise_std, mean, std, stream):
    """Marshal one bseg_train_aug_fwd call (fn = the C entry point).  Returns (out_image, out_mask, colour)."""
    B, _, H, W = image.shape
    out = torch.empty_like(image)
    colour = torch.empty_like(image)
    out_mask = torch.empty_like(mask) if mask is not None else None
    rc = fn(_lib.ptr(image), _lib.ptr(mask), _lib.ptr(params), _order_arr(order), _lib.ptr(noise), float(noise_mean),
            float(noise_std), _lib.f3(mean), _lib.f3(std), _lib.ptr(out), _lib.ptr(out_mask), _lib.ptr(colour), B, H, W,
            *stream)
    return rc, out, out_mask, colour


def _raw_bwd(fn, image, params, order, std, colour, d_out, stream):
    B, _, H, W = image.shape
    scratch = torch.empty((2,) + tuple(image.shape), dtype=torch.float32, device=image.device)
    d_image = torch.empty_like(image)
    rc = fn(_lib.ptr(image), _lib.ptr(params), _order_arr(order), _lib.f3(std), _lib.ptr(colour), _lib.ptr(d_out),
            _lib.ptr(scratch), _lib.ptr(d_image), B, H, W, *stream)
    return rc, d_image


class _TrainAugFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, mask, params, order, noise, noise_mean, noise_std, mean, std):
        if not image.is_cuda:
            raise _lib.BsegError("train_aug: tensors must be on a CUDA device (there is no CPU fallback)")
        image = image.detach().contiguous().float()
        with torch.cuda.device(image.device):
            rc, out, out_mask, colour = _raw_fwd(_lib.lib().bseg_train_aug_fwd, image, mask, params, order, noise,
                                                 noise_mean, noise_std, mean, std, (_lib.stream_ptr(),))
        _lib.check(rc, "bseg_train_aug_fwd")
        ctx.save_for_backward(image, params, colour)
        ctx.order, ctx.std = tuple(order), tuple(std)
        if out_mask is None:
            out_mask = torch.empty(0, dtype=torch.uint8, device=image.device)
        ctx.mark_non_differentiable(out_mask)
        return out, out_mask

    @staticmethod
    def backward(ctx, d_out, _d_mask):
        image, params, colour = ctx.saved_tensors
        with torch.cuda.device(image.device):
            rc, d_image = _raw_bwd(_lib.lib().bseg_train_aug_bwd, image, params, ctx.order, ctx.std, colour,
                                   d_out.contiguous().float(), (_lib.stream_ptr(),))
        _lib.check(rc, "bseg_train_aug_bwd")
        return d_image, None, None, None, None, None, None, None, None


def pack_params(B: int, *, vflip=None, hflip=None, brightness=None, contrast=None, saturation=None, hue=None,
                sharp_apply=None, sharp_factor=None, erase_apply=None, erase_box=None, erase_value: float = 0.0,
                noise_apply=None) -> torch.Tensor:
    """The [B,16] float32 parameter rows of include/bseg.h from per-sample draws (CPU tensors; identity where None).
    `brightness`, `contrast`, `saturation` are kornia's factors (identity 1), `hue` is in turns (identity 0)."""
    f32 = torch.float32
    P = torch.zeros((B, N_PARAMS), dtype=f32)
    as_f = lambda t, default: (torch.full((B,), default, dtype=f32) if t is None else torch.as_tensor(t).to(f32))
    P[:, P_VFLIP] = as_f(vflip, 0.0)
    P[:, P_HFLIP] = as_f(hflip, 0.0)
    P[:, P_BRIGHT] = as_f(brightness, 1.0) - 1            # ColorJiggle: adjust_brightness(x, factor - 1)
    P[:, P_CONTRAST] = as_f(contrast, 1.0)
    P[:, P_SATURATION] = as_f(saturation, 1.0)
    P[:, P_HUE_RAD] = as_f(hue, 0.0) * 2 * math.pi        # ColorJiggle: adjust_hue(x, factor * 2 * pi)
    P[:, P_SHARP_ON] = as_f(sharp_apply, 0.0)
    P[:, P_SHARP_F] = as_f(sharp_factor, 1.0)
    P[:, P_ERASE_ON] = as_f(erase_apply, 0.0)
    if erase_box is not None:
        P[:, P_ERASE_X:P_ERASE_H + 1] = torch.as_tensor(erase_box).to(f32)
    P[:, P_ERASE_VALUE] = erase_value
    P[:, P_NOISE_ON] = as_f(noise_apply, 0.0)
    return P


class TrainAug:
    """src/data.py:195-224.  `generator`: a CPU torch.Generator for the parameter draws (default: the global one); the
    Gaussian noise field is drawn on the device."""

    def __init__(self, conf, mean=IMAGE_MEAN, std=IMAGE_STD, generator: Optional[torch.Generator] = None,
                 erase_ratio=(0.3, 3.3), erase_value: float = 0.0):
        self.conf = conf
        self.mean, self.std = tuple(mean), tuple(std)
        self.generator = generator
        self.erase_ratio = erase_ratio
        self.erase_value = erase_value
        self.last_params: Optional[dict] = None

    def __len__(self):
        return 7

    # ---- parameter draws (kornia/augmentation/random_generator/_2d/{color_jiggle,rectangle_earse,plain_uniform}.py) ----
    def _u(self, n, lo, hi):
        return torch.rand(n, generator=self.generator) * (hi - lo) + lo

    def _bern(self, n, p):
        return torch.rand(n, generator=self.generator) < p

    def sample_params(self, B: int, H: int, W: int) -> dict:
        c = self.conf
        clip = lambda lo, hi, a, b: (max(lo, a), min(hi, b))
        b_lo, b_hi = clip(1 - c.brightness, 1 + c.brightness, 0.0, 2.0)
        c_lo, c_hi = clip(1 - c.contrast, 1 + c.contrast, 0.0, float("inf"))
        s_lo, s_hi = clip(1 - c.saturation, 1 + c.saturation, 0.0, float("inf"))
        h_lo, h_hi = clip(-c.hue, c.hue, -0.5, 0.5)
        d = {
            "vflip": self._bern(B, c.vertical_flip), "hflip": self._bern(B, c.horizontal_flip),
            "brightness": self._u(B, b_lo, b_hi), "contrast": self._u(B, c_lo, c_hi),
            "saturation": self._u(B, s_lo, s_hi), "hue": self._u(B, h_lo, h_hi),
            "order": tuple(torch.randperm(4, generator=self.generator).tolist()),
            "sharp_apply": self._bern(B, c.sharpness_p), "sharp_factor": self._u(B, 0.0, max(c.sharpness, 0.0)),
            "erase_apply": self._bern(B, c.erasing_p), "noise_apply": self._bern(B, c.gauss_p),
        }
        # RectangleEraseGenerator: area fraction ~ U(scale); aspect ratio ~ U(r0,1) or U(1,r1) with equal probability
        # when r0 < 1 < r1, else U(r0,r1); height = round(sqrt(area*ar)), width = round(sqrt(area/ar)), clipped to
        # [1, size]; top-left corner uniform over the positions that keep the box inside
        area = self._u(B, c.erasing_scale[0], c.erasing_scale[1]) * (H * W)
        r0, r1 = self.erase_ratio
        if r0 < 1.0 < r1:
            ar = torch.where(self._bern(B, 0.5), self._u(B, r0, 1.0), self._u(B, 1.0, r1))
        else:
            ar = self._u(B, r0, r1)
        eh = torch.clamp(torch.round(torch.sqrt(area * ar)), 1, H)
        ew = torch.clamp(torch.round(torch.sqrt(area / ar)), 1, W)
        ex = torch.floor(self._u(B, 0.0, 1.0) * (W - ew + 1))
        ey = torch.floor(self._u(B, 0.0, 1.0) * (H - eh + 1))
        d["erase_box"] = torch.stack([ex, ey, ew, eh], dim=1).to(torch.int64)
        return d

    def apply(self, image: torch.Tensor, mask: Optional[torch.Tensor], d: dict, noise: Optional[torch.Tensor] = None):
        """The chain for one explicit parameter draw `d` (keys of `sample_params`).  Returns (image, mask)."""
        B, _, H, W = image.shape
        params = pack_params(B, vflip=d["vflip"], hflip=d["hflip"], brightness=d["brightness"], contrast=d["contrast"],
                             saturation=d["saturation"], hue=d["hue"], sharp_apply=d["sharp_apply"],
                             sharp_factor=d["sharp_factor"], erase_apply=d["erase_apply"], erase_box=d["erase_box"],
                             erase_value=self.erase_value, noise_apply=d["noise_apply"]).to(image.device)
        if noise is None and bool(torch.as_tensor(d["noise_apply"]).any()):
            noise = torch.randn(image.shape, dtype=torch.float32, device=image.device)
        m8 = None
        if mask is not None:
            m8 = mask.reshape(B, H, W).to(torch.uint8).contiguous()
        out, out_mask = _TrainAugFn.apply(image, m8, params, d["order"], noise, self.conf.gauss_mean,
                                          self.conf.gauss_std, self.mean, self.std)
        if mask is None:
            return out, None
        return out, out_mask.to(mask.dtype).reshape(mask.shape)

    def __call__(self, batch: dict) -> dict:
        image = batch["image"]
        B, _, H, W = image.shape
        d = self.sample_params(B, H, W)
        self.last_params = d
        out = dict(batch)
        out["image"], m = self.apply(image, batch.get("mask"), d)
        if m is not None:
            out["mask"] = m
        return out


class InferenceAug:
    """Stand-in for the reference datamodule's `aug` pipeline (src/data.py:226-234): CenterCrop(inpt_size) is the
    identity on inpt_size inputs, Normalize(mean, std) is applied to the image entry; masks pass through."""

    def __call__(self, batch: dict) -> dict:
        out = dict(batch)
        mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32, device=batch["image"].device).view(1, 3, 1, 1)
        std = torch.tensor(IMAGE_STD, dtype=torch.float32, device=batch["image"].device).view(1, 3, 1, 1)
        out["image"] = (batch["image"] - mean) / std
        return out

    def __len__(self):
        return 2


class Augmentations:
    """The two attributes `PromptModel.post_init(datamodule)` reads from the reference's datamodule
    (src/model.py:104-106): `train_aug` and `aug`."""

    def __init__(self, conf, generator: Optional[torch.Generator] = None):
        self.train_aug = TrainAug(conf, generator=generator)
        self.aug = InferenceAug()
