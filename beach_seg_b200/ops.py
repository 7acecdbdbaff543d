"""Python wrappers (device tensors in, device tensors out) over the C-ABI kernels that replace the tensor glue of
the reference around the SegGPT call.  Every function names the reference lines it stands in for.  No CPU fallback:
tensors must live on a CUDA device."""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib

IMAGE_MEAN = (0.485, 0.456, 0.406)  # SegGptImageProcessor.image_mean (HF:image_processing_seggpt.py:76)
IMAGE_STD = (0.229, 0.224, 0.225)   # SegGptImageProcessor.image_std  (HF:image_processing_seggpt.py:77)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.BsegError("beach_seg_b200.ops works on CUDA tensors only (no CPU fallback)")


# ------------------------------------------------------------------------------------------------------------
# ingest
# ------------------------------------------------------------------------------------------------------------
def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


@lru_cache(maxsize=16)
def pil_bicubic_table(in_size: int, out_size: int = 448):
    """Coefficient table of PIL's 8-bit BICUBIC resampler (what `Image.resize(..., BICUBIC)` at src/data.py:93-96
    precomputes): bounds int32 [out,2] = (first tap, tap count), coef int32 [out,ksize] in 22-bit fixed point.
    in_size == out_size yields the identity table (the reference skips the resize in that case)."""
    if in_size == out_size:
        bounds = np.stack([np.arange(out_size), np.ones(out_size)], axis=1).astype(np.int32)
        coef = np.full((out_size, 1), 1 << 22, dtype=np.int32)
        return bounds, coef
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coef = np.zeros((out_size, ksize), dtype=np.int32)
    inv = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [_bicubic((x + xmin - center + 0.5) * inv) for x in range(xmax)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            coef[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, coef


_table_cache: dict = {}


def _device_table(in_size: int, device):
    key = (in_size, str(device))
    if key not in _table_cache:
        bounds, coef = pil_bicubic_table(in_size, 448)
        _table_cache[key] = (torch.from_numpy(bounds).to(device), torch.from_numpy(coef).to(device), coef.shape[1])
    return _table_cache[key]


def scene_stats(scene_u16: torch.Tensor, nodata: torch.Tensor) -> torch.Tensor:
    """Scene-global statistics of tif_image's 4-band branch (src/util/geo_util.py:459-464).
    scene_u16: uint16 [4,Hs,Ws] (torch.uint16 or int16 storage); nodata: bool/uint8 [Hs,Ws].
    Returns float32 [4] = (min over valid composite pixels, max of channel 0, 1, 2)."""
    _need_cuda(scene_u16, nodata)
    _, Hs, Ws = scene_u16.shape
    nd = nodata.to(torch.uint8).contiguous()
    stats = torch.empty(4, dtype=torch.float32, device=scene_u16.device)
    scratch = torch.empty(4, dtype=torch.int32, device=scene_u16.device)
    with torch.cuda.device(scene_u16.device):
        _lib.check(_lib.lib().bseg_scene_stats(_lib.ptr(scene_u16), _lib.ptr(nd), Hs, Ws, _lib.ptr(stats),
                                               _lib.ptr(scratch), _lib.stream_ptr()), "bseg_scene_stats")
    return stats


def ingest_tiles(scene_u16: torch.Tensor, nodata: torch.Tensor, stats: torch.Tensor, boxes: torch.Tensor,
                 crop: int, want_nchw: bool = True, want_u8: bool = False, want_nodata: bool = False,
                 out_patch: Optional[torch.Tensor] = None, patch_tile_stride: int = 0):
    """tif_image + crop_tif + PIL BICUBIC resize to 448 + /255 + Normalize for a batch of tile boxes
    (src/util/geo_util.py:454-468,297-341; src/data.py:93-124,226-229).
    boxes: int32 [n,4] (xmin,ymin,xmax,ymax) on the device.  Returns dict with the requested outputs:
    image float32 [n,3,448,448], u8 uint8 [n,crop,crop,3], nodata uint8 [n,crop,crop]."""
    _need_cuda(scene_u16, nodata, stats, boxes)
    dev = scene_u16.device
    _, Hs, Ws = scene_u16.shape
    n = boxes.shape[0]
    nd = nodata.to(torch.uint8).contiguous()
    bounds, coef, ksize = _device_table(crop, dev)
    out = {}
    out["image"] = torch.empty((n, 3, 448, 448), dtype=torch.float32, device=dev) if want_nchw else None
    out["u8"] = torch.empty((n, crop, crop, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    out["nodata"] = torch.empty((n, crop, crop), dtype=torch.uint8, device=dev) if want_nodata else None
    boxes_i = boxes.to(torch.int32).contiguous()  # named: must outlive the launch
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_ingest_u16x4(
            _lib.ptr(scene_u16), _lib.ptr(nd), Hs, Ws, _lib.ptr(stats), _lib.ptr(boxes_i),
            n, crop, _lib.ptr(coef), _lib.ptr(bounds), ksize, _lib.f3(IMAGE_MEAN), _lib.f3(IMAGE_STD),
            _lib.ptr(out["image"]), _lib.ptr(out_patch), patch_tile_stride, _lib.ptr(out["u8"]),
            _lib.ptr(out["nodata"]), _lib.stream_ptr()), "bseg_ingest_u16x4")
    return out


# ------------------------------------------------------------------------------------------------------------
# palettes / colourise / decode
# ------------------------------------------------------------------------------------------------------------
def colorize_norm(mask: torch.Tensor, palette: torch.Tensor) -> torch.Tensor:
    """normalize(torch_apply_mask_rgb(palette, mask)) (src/util/ml_util.py:114-132; src/model.py:210-211,238-239).
    mask: integer [B,1,H,W] or [B,H,W]; palette: uint8 [B,C,3].  Returns float32 [B,3,H,W]."""
    _need_cuda(mask, palette)
    if mask.ndim == 4:
        mask = mask.squeeze(1)
    B, H, W = mask.shape
    m8 = mask.to(torch.uint8).contiguous()
    pal = palette.to(torch.uint8).contiguous()
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=mask.device)
    with torch.cuda.device(mask.device):
        _lib.check(_lib.lib().bseg_colorize_norm(_lib.ptr(m8), _lib.ptr(pal), pal.shape[1], _lib.f3(IMAGE_MEAN),
                                                 _lib.f3(IMAGE_STD), _lib.ptr(out), B, H, W, _lib.stream_ptr()),
                   "bseg_colorize_norm")
    return out


@lru_cache(maxsize=16)
def cv2_nearest_index(src: int, dst: int) -> np.ndarray:
    """Source index per destination index of cv2.resize(..., INTER_NEAREST) (src/predict.py:258)."""
    ifx = 1.0 / (dst / src)
    return np.array([min(int(math.floor(x * ifx)), src - 1) for x in range(dst)], dtype=np.int32)


def decode_palette(pred_masks: torch.Tensor, palette_norm: torch.Tensor, out_size: Optional[int] = None,
                   nodata: Optional[torch.Tensor] = None, dtype=torch.int64) -> torch.Tensor:
    """PromptModel.process_pred_masks (src/model.py:155-175), optionally fused with the cv2 INTER_NEAREST resize
    back to crop size (src/predict.py:258) and nodata zeroing (src/predict_no_prompt.py:303).
    pred_masks: float32 [B,3,2H,W]; palette_norm: float32 [B,C,3]. Returns [B,out,out] int64 (or uint8)."""
    _need_cuda(pred_masks, palette_norm, nodata)
    B, _, H2, W = pred_masks.shape
    H = H2 // 2
    out_size = H if out_size is None else int(out_size)
    dev = pred_masks.device
    idx = None
    if out_size != H or out_size != W:
        idx = torch.from_numpy(cv2_nearest_index(H, out_size)).to(dev)
    o8 = torch.empty((B, out_size, out_size), dtype=torch.uint8, device=dev) if dtype == torch.uint8 else None
    o64 = torch.empty((B, out_size, out_size), dtype=torch.int64, device=dev) if dtype == torch.int64 else None
    nd = nodata.to(torch.uint8).contiguous() if nodata is not None else None
    pred_c, pal_c = pred_masks.contiguous(), palette_norm.to(torch.float32).contiguous()  # must outlive the launch
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_decode_palette(
            _lib.ptr(pred_c), _lib.ptr(pal_c),
            palette_norm.shape[1], _lib.ptr(o8), _lib.ptr(o64), _lib.ptr(nd), _lib.ptr(idx), B, H, W, out_size,
            _lib.stream_ptr()), "bseg_decode_palette")
    return o8 if o8 is not None else o64


def mean_over_prompts(pred_masks: torch.Tensor, prompts: int) -> torch.Tensor:
    """pred_masks.mean(dim=0, keepdim=True) per tile of `prompts` samples (src/predict_no_prompt.py:298)."""
    _need_cuda(pred_masks)
    B = pred_masks.shape[0]
    n_tiles = B // prompts
    per = pred_masks[0].numel()
    out = torch.empty((n_tiles, *pred_masks.shape[1:]), dtype=torch.float32, device=pred_masks.device)
    pred_c = pred_masks.contiguous()
    with torch.cuda.device(pred_masks.device):
        _lib.check(_lib.lib().bseg_mean_over_prompts(_lib.ptr(pred_c), _lib.ptr(out), n_tiles,
                                                     prompts, per, _lib.stream_ptr()), "bseg_mean_over_prompts")
    return out


# ------------------------------------------------------------------------------------------------------------
# vote stitching
# ------------------------------------------------------------------------------------------------------------
def vote_accumulate(counter: torch.Tensor, cls: torch.Tensor, boxes: torch.Tensor, overlapping: bool = True) -> None:
    """Accumulator.update for a batch of tiles (src/predict.py:120-159).  counter: int32 [Hs,Ws] whose bytes are the
    reference's uint8 (Hs,Ws,4) vote counters; cls: uint8 [n,crop,crop]; boxes: int32 [n,4]."""
    _need_cuda(counter, cls, boxes)
    Hs, Ws = counter.shape
    n, crop, _ = cls.shape
    cls_c, boxes_i = cls.contiguous(), boxes.to(torch.int32).contiguous()  # must outlive the launch
    with torch.cuda.device(counter.device):
        _lib.check(_lib.lib().bseg_vote_accumulate(_lib.ptr(counter), Hs, Ws, _lib.ptr(cls_c), n, crop,
                                                   _lib.ptr(boxes_i),
                                                   1 if (overlapping and n > 1) else 0, _lib.stream_ptr()),
                   "bseg_vote_accumulate")


def vote_argmax(counter: torch.Tensor) -> torch.Tensor:
    """np.argmax(counter, axis=2) (src/predict.py:100) -> uint8 [Hs,Ws]."""
    _need_cuda(counter)
    out = torch.empty(counter.shape, dtype=torch.uint8, device=counter.device)
    with torch.cuda.device(counter.device):
        _lib.check(_lib.lib().bseg_vote_argmax(_lib.ptr(counter), _lib.ptr(out), counter.numel(), _lib.stream_ptr()),
                   "bseg_vote_argmax")
    return out


# ------------------------------------------------------------------------------------------------------------
# loss
# ------------------------------------------------------------------------------------------------------------
def smooth_l1_loss(pred_masks: torch.Tensor, labels: torch.Tensor, yesdata: torch.Tensor, beta: float,
                   per_sample: bool = False, want_grad: bool = False):
    """SegGptLoss.forward of the reference (src/model.py:45-64) and d(loss)/d(pred_masks).
    per_sample=False is the code as written (BxB keep-mask broadcast at B>1)."""
    _need_cuda(pred_masks, labels, yesdata)
    B, _, H2, W = pred_masks.shape
    H = H2 // 2
    dev = pred_masks.device
    yes = yesdata.reshape(B, H, W).to(torch.uint8).contiguous()
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    grad = torch.empty_like(pred_masks, dtype=torch.float32) if want_grad else None
    scratch = torch.empty(2, dtype=torch.float32, device=dev)
    pred_c, lab_c = pred_masks.contiguous(), labels.to(torch.float32).contiguous()  # must outlive the launch
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_loss_smoothl1_fwd_bwd(
            _lib.ptr(pred_c), _lib.ptr(lab_c), _lib.ptr(yes),
            float(beta), 1 if per_sample else 0, _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(scratch), B, H, W,
            _lib.stream_ptr()), "bseg_loss_smoothl1_fwd_bwd")
    return (loss[0], grad) if want_grad else loss[0]
