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
LOSS_SCRATCH_FLOATS = 2052            # BSEG_LOSS_SCRATCH_FLOATS (include/bseg.h)


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
_index_cache: dict = {}


def _device_array(kind: str, key, device, make):
    """Device copy of a small host array, cached per (kind, key, device): a fresh `torch.from_numpy(...).to(dev)` is a
    blocking pageable copy that synchronises the stream on every call (and stalls the upload/compute overlap of
    predict.HostScenePipeline)."""
    k = (kind, key, str(device))
    t = _index_cache.get(k)
    if t is None:
        t = torch.from_numpy(np.ascontiguousarray(make())).to(device)
        _index_cache[k] = t
    return t


def _device_table(in_size: int, device):
    key = (in_size, str(device))
    if key not in _table_cache:
        bounds, coef = pil_bicubic_table(in_size, 448)
        _table_cache[key] = (torch.from_numpy(bounds).to(device), torch.from_numpy(coef).to(device), coef.shape[1])
    return _table_cache[key]


def scene_stats(scene_u16: torch.Tensor, nodata: torch.Tensor) -> torch.Tensor:
    """Scene-global statistics of tif_image's 4-band branch (src/util/geo_util.py:459-464).
    scene_u16: uint16 [4,Hs,Ws] (torch.uint16 or int16 storage) or float32 [4,Hs,Ws] (the merge_tifs mosaic,
    src/util/geo_util.py:385,417); nodata: bool/uint8 [Hs,Ws].
    Returns float32 [4] = (min over valid composite pixels, max of channel 0, 1, 2)."""
    _need_cuda(scene_u16, nodata)
    scene_u16, is_f32 = _scene_kind(scene_u16)
    _, Hs, Ws = scene_u16.shape
    nd = nodata.to(torch.uint8).contiguous()
    stats = torch.empty(4, dtype=torch.float32, device=scene_u16.device)
    scratch = torch.empty(4, dtype=torch.int32, device=scene_u16.device)
    with torch.cuda.device(scene_u16.device):
        fn = _lib.lib().bseg_scene_stats_f32 if is_f32 else _lib.lib().bseg_scene_stats
        _lib.check(fn(_lib.ptr(scene_u16), _lib.ptr(nd), Hs, Ws, _lib.ptr(stats), _lib.ptr(scratch),
                      _lib.stream_ptr()), "bseg_scene_stats")
    return stats


def scene_stats_sharded(scene_u16: torch.Tensor, nodata: torch.Tensor, row0: int, row1: int, group=None) -> torch.Tensor:
    """`scene_stats` of a scene whose rows are spread over the ranks of `group`: this rank reduces rows [row0, row1) of
    its device copy (rows outside are never read, so they need not have been uploaded), the ranks' order-preserving
    integer keys are merged with one MAX all-reduce (the min key negated), and the merged keys are decoded.  The row
    ranges of the ranks must cover the scene; the result is bit-identical to `scene_stats` of the whole scene."""
    _need_cuda(scene_u16, nodata)
    scene_u16, is_f32 = _scene_kind(scene_u16)
    _, Hs, Ws = scene_u16.shape
    nd = nodata.to(torch.uint8).contiguous()
    dev = scene_u16.device
    keys = torch.empty(4, dtype=torch.int32, device=dev)
    stats = torch.empty(4, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_scene_stats_rows(_lib.ptr(scene_u16), int(is_f32), _lib.ptr(nd), Hs, Ws, int(row0),
                                                    int(row1), _lib.ptr(keys), _lib.stream_ptr()),
                   "bseg_scene_stats_rows")
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            k = keys.to(torch.int64) & 0xFFFFFFFF            # the uint32 keys
            k[0] = -k[0]                                       # min -> max
            dist.all_reduce(k, op=dist.ReduceOp.MAX, group=group)
            k[0] = -k[0]
            keys = k.to(torch.int32)                           # keeps the low 32 bits
        _lib.check(_lib.lib().bseg_scene_stats_finalize(_lib.ptr(keys), _lib.ptr(stats), _lib.stream_ptr()),
                   "bseg_scene_stats_finalize")
    return stats


def _scene_kind(scene: torch.Tensor):
    """(contiguous scene, is_float32).  uint16 / int16 storage -> the u16 entry points, float32 -> the f32 ones."""
    if scene.dtype == torch.float32:
        return scene.contiguous(), True
    if scene.dtype in (torch.uint16, torch.int16):
        return scene.contiguous(), False
    raise _lib.BsegError(f"scene must be uint16 or float32 [4,Hs,Ws], got {scene.dtype}")


def merge_mosaic(data: torch.Tensor, yesdata: torch.Tensor):
    """The accumulation of merge_tifs (src/util/geo_util.py:410-419) for rasters already on the output grid.
    data: float32 [N,C,H,W]; yesdata: uint8 [N,H,W] (rasterio masks: 0 / 255, used as weights like the reference).
    Returns (mean float32 [C,H,W], nodata bool [H,W])."""
    _need_cuda(data, yesdata)
    if data.dtype != torch.float32 or yesdata.dtype != torch.uint8:
        raise _lib.BsegError("merge_mosaic: data must be float32 and yesdata uint8")
    N, C, H, W = data.shape
    if tuple(yesdata.shape) != (N, H, W):
        raise _lib.BsegError(f"merge_mosaic: yesdata shape {tuple(yesdata.shape)} != {(N, H, W)}")
    d = data.contiguous()
    y = yesdata.contiguous()
    mean = torch.empty((C, H, W), dtype=torch.float32, device=data.device)
    nodata = torch.empty((H, W), dtype=torch.uint8, device=data.device)
    with torch.cuda.device(data.device):
        _lib.check(_lib.lib().bseg_merge_mosaic(_lib.ptr(d), _lib.ptr(y), N, C, H, W, _lib.ptr(mean),
                                                _lib.ptr(nodata), _lib.stream_ptr()), "bseg_merge_mosaic")
    return mean, nodata.bool()


def ingest_tiles(scene_u16: torch.Tensor, nodata: torch.Tensor, stats: torch.Tensor, boxes: torch.Tensor,
                 crop: int, want_nchw: bool = True, want_u8: bool = False, want_nodata: bool = False,
                 out_patch: Optional[torch.Tensor] = None, patch_tile_stride: int = 0, normalize: bool = True,
                 out_size: int = 448):
    """tif_image + crop_tif + PIL BICUBIC resize to 448 + /255 + Normalize for a batch of tile boxes
    (src/util/geo_util.py:454-468,297-341; src/data.py:93-124,226-229).
    boxes: int32 [n,4] (xmin,ymin,xmax,ymax) on the device.  Returns dict with the requested outputs:
    image float32 [n,3,448,448], u8 uint8 [n,crop,crop,3], nodata uint8 [n,crop,crop].
    normalize=False stops after `/255` (the dataset item of src/data.py:93-96, before any augmentation pipeline): the
    input of `augment.TrainAug` for the training batch (src/data.py:295-313), which ends with Normalize itself.
    out_size: 448 (the reference's inpt_size) or == crop for the native-resolution mode, where get_crop skips the
    resize (src/data.py:94) and the image is float32 [n,3,crop,crop]."""
    _need_cuda(scene_u16, nodata, stats, boxes)
    scene_u16, is_f32 = _scene_kind(scene_u16)
    dev = scene_u16.device
    _, Hs, Ws = scene_u16.shape
    n = boxes.shape[0]
    nd = nodata.to(torch.uint8).contiguous()
    if out_size != 448:
        if out_size != crop:
            raise _lib.BsegError(f"ingest_tiles: out_size must be 448 or the crop size ({crop}), got {out_size}")
        if out_patch is not None or not want_nchw:
            raise _lib.BsegError("ingest_tiles: the native-resolution path writes the NCHW image only")
        out = {"image": torch.empty((n, 3, crop, crop), dtype=torch.float32, device=dev),
               "u8": torch.empty((n, crop, crop, 3), dtype=torch.uint8, device=dev) if want_u8 else None,
               "nodata": torch.empty((n, crop, crop), dtype=torch.uint8, device=dev) if want_nodata else None}
        boxes_n = boxes.to(torch.int32).contiguous()
        with torch.cuda.device(dev):
            fn = _lib.lib().bseg_ingest_native_f32x4 if is_f32 else _lib.lib().bseg_ingest_native_u16x4
            _lib.check(fn(_lib.ptr(scene_u16), _lib.ptr(nd), Hs, Ws, _lib.ptr(stats), _lib.ptr(boxes_n), n, crop,
                          _lib.f3(IMAGE_MEAN if normalize else (0.0, 0.0, 0.0)),
                          _lib.f3(IMAGE_STD if normalize else (1.0, 1.0, 1.0)), _lib.ptr(out["image"]),
                          _lib.ptr(out["u8"]), _lib.ptr(out["nodata"]), _lib.stream_ptr()), "bseg_ingest_native")
        return out
    bounds, coef, ksize = _device_table(crop, dev)
    out = {}
    out["image"] = torch.empty((n, 3, 448, 448), dtype=torch.float32, device=dev) if want_nchw else None
    out["u8"] = torch.empty((n, crop, crop, 3), dtype=torch.uint8, device=dev) if want_u8 else None
    out["nodata"] = torch.empty((n, crop, crop), dtype=torch.uint8, device=dev) if want_nodata else None
    boxes_i = boxes.to(torch.int32).contiguous()  # named: must outlive the launch
    with torch.cuda.device(dev):
        fn = _lib.lib().bseg_ingest_f32x4 if is_f32 else _lib.lib().bseg_ingest_u16x4
        _lib.check(fn(
            _lib.ptr(scene_u16), _lib.ptr(nd), Hs, Ws, _lib.ptr(stats), _lib.ptr(boxes_i),
            n, crop, _lib.ptr(coef), _lib.ptr(bounds), ksize, _lib.f3(IMAGE_MEAN if normalize else (0.0, 0.0, 0.0)),
            _lib.f3(IMAGE_STD if normalize else (1.0, 1.0, 1.0)),
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
        idx = _device_array("cv2_nearest", (H, out_size), dev, lambda: cv2_nearest_index(H, out_size))
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


CLASS_COLORS = {"nodata": None, "water": "yellow", "veg": "blue", "sand": "hotpink"}
"""src/util/img_util.py:12."""
_COLOR_RGB = {"yellow": (255, 255, 0), "blue": (0, 0, 255), "hotpink": (255, 105, 180)}  # ImageColor.getrgb values


def class_rgba_table(classes: Sequence[str], alpha: int = int(255 * 0.3)) -> np.ndarray:
    """uint8 [n_classes,4] overlay table of overlay_prediction (src/util/img_util.py:105-111): (r,g,b,alpha) per class
    id, alpha 0 for classes whose CLASS_COLORS entry is None."""
    table = np.zeros((len(classes), 4), dtype=np.uint8)
    for i, c in enumerate(classes):
        name = CLASS_COLORS[c]
        if name is not None:
            table[i] = (*_COLOR_RGB[name], alpha)
    return table


def paste_tiles(canvas: torch.Tensor, crops_u8: torch.Tensor, boxes: torch.Tensor) -> None:
    """Accumulator.update's `current_img[dy0:dy1, dx0:dx1] = img_crop[sy0:sy1, sx0:sx1]` (src/predict.py:157) for a batch
    of tiles.  canvas: uint8 [Hs,Ws,3] (modified in place); crops_u8: uint8 [n,crop,crop,3]; boxes: int32 [n,4]."""
    _need_cuda(canvas, crops_u8, boxes)
    if canvas.dtype != torch.uint8 or crops_u8.dtype != torch.uint8 or not canvas.is_contiguous():
        raise _lib.BsegError("paste_tiles: canvas and crops must be uint8, canvas contiguous")
    Hs, Ws, _ = canvas.shape
    n, crop = crops_u8.shape[0], crops_u8.shape[1]
    cr = crops_u8.contiguous()
    bx = boxes.to(torch.int32).contiguous()
    with torch.cuda.device(canvas.device):
        _lib.check(_lib.lib().bseg_paste_tiles_u8(_lib.ptr(canvas), Hs, Ws, _lib.ptr(cr), n, crop, _lib.ptr(bx),
                                                  _lib.stream_ptr()), "bseg_paste_tiles_u8")


def overlay_prediction(img: torch.Tensor, pred: torch.Tensor, classes: Sequence[str]) -> torch.Tensor:
    """overlay_prediction (src/util/img_util.py:98-116) on the device, bit-exact with Pillow's alpha_composite.
    img: uint8 [H,W,3]; pred: uint8 [H,W] class ids.  Returns uint8 [H,W,3]."""
    _need_cuda(img, pred)
    H, W, _ = img.shape
    im = img.to(torch.uint8).contiguous()
    pr = pred.to(torch.uint8).contiguous()
    table = _device_array("class_rgba", tuple(classes), img.device, lambda: class_rgba_table(classes))
    out = torch.empty_like(im)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().bseg_overlay_prediction(_lib.ptr(im), _lib.ptr(pr), _lib.ptr(table), len(classes),
                                                      H * W, _lib.ptr(out), _lib.stream_ptr()),
                   "bseg_overlay_prediction")
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
    scratch = torch.empty(LOSS_SCRATCH_FLOATS, dtype=torch.float32, device=dev)
    pred_c, lab_c = pred_masks.contiguous(), labels.to(torch.float32).contiguous()  # must outlive the launch
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_loss_smoothl1_fwd_bwd(
            _lib.ptr(pred_c), _lib.ptr(lab_c), _lib.ptr(yes),
            float(beta), 1 if per_sample else 0, _lib.ptr(loss), _lib.ptr(grad), _lib.ptr(scratch), B, H, W,
            _lib.stream_ptr()), "bseg_loss_smoothl1_fwd_bwd")
    return (loss[0], grad) if want_grad else loss[0]


# ------------------------------------------------------------------------------------------------------------
# HF image-processor path (src/predict_no_prompt.py:235-301; HF:image_processing_seggpt.py)
# ------------------------------------------------------------------------------------------------------------
@lru_cache(maxsize=16)
def tv_bicubic_aa_table(in_size: int, out_size: int = 448):
    """Coefficient table of torchvision's uint8 bicubic-antialias resize, the kernel the HF processor's
    `tvF.resize(..., BICUBIC, antialias=True)` runs on CPU (HF:image_processing_backends.py:200-251 -> ATen
    upsample_avx_bilinear_bicubic_uint8 / _compute_index_ranges_int16_weights): Pillow's windowed bicubic (a = -0.5,
    support 2*scale) with weights quantised to int16 at the largest precision that keeps them below 2^15.
    Returns (bounds int32 [out,2], coef int32 [out,ksize], precision_bits).  Bit-exact against torchvision
    (tests/test_processor.py)."""
    if in_size == out_size:
        bounds = np.stack([np.arange(out_size), np.ones(out_size)], axis=1).astype(np.int32)
        return bounds, np.full((out_size, 1), 1 << 14, dtype=np.int32), 14
    scale = in_size / out_size
    support = 2.0 * scale if scale >= 1.0 else 2.0
    ksize = int(math.ceil(support)) * 2 + 1
    inv = 1.0 / scale if scale >= 1.0 else 1.0
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    w = np.zeros((out_size, ksize), dtype=np.float64)
    wmax = 0.0
    for i in range(out_size):
        center = scale * (i + 0.5)
        xmin = max(int(center - support + 0.5), 0)
        xsize = min(max(min(int(center + support + 0.5), in_size) - xmin, 0), ksize)
        ws = [_bicubic((j + xmin - center + 0.5) * inv) for j in range(xsize)]
        tot = sum(ws)
        if tot != 0.0:
            ws = [v / tot for v in ws]
        wmax = max([wmax] + ws)
        w[i, :xsize] = ws
        bounds[i] = (xmin, xsize)
    prec = 0
    while prec < 22 and int(0.5 + wmax * (1 << (prec + 1))) < (1 << 15):
        prec += 1
    v = w * (1 << prec)
    coef = np.where(v < 0, np.trunc(v - 0.5), np.trunc(v + 0.5)).astype(np.int32)
    return bounds, coef, prec


@lru_cache(maxsize=16)
def torch_nearest_index(src: int, dst: int) -> np.ndarray:
    """Source index per destination index of torch `interpolate(mode="nearest")` / torchvision NEAREST resize:
    min(floor(dst_index * float32(src / dst)), src - 1), evaluated in float32 like ATen."""
    scale = np.float32(src) / np.float32(dst)
    idx = np.floor(np.arange(dst, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(idx, src - 1).astype(np.int32)


@lru_cache(maxsize=16)
def torch_nearest_exact_index(src: int, dst: int) -> np.ndarray:
    """The same for `mode="nearest-exact"` (torchvision NEAREST_EXACT, what HF maps PIL NEAREST to,
    transformers/image_utils.py:56): min(floor((dst_index + 0.5) * float32(src / dst)), src - 1)."""
    scale = np.float32(src) / np.float32(dst)
    idx = np.floor((np.arange(dst, dtype=np.float32) + np.float32(0.5)) * scale).astype(np.int64)
    return np.minimum(idx, src - 1).astype(np.int32)


def _hf_mean_std_255():
    """HF's fused rescale+normalise constants: tensor(mean) * (1.0 / rescale_factor), rescale_factor = 1/255
    (HF:image_processing_backends.py:292-306), in float32."""
    f = 1.0 / (1 / 255)
    return ((torch.tensor(IMAGE_MEAN) * f).tolist(), (torch.tensor(IMAGE_STD) * f).tolist())


_tv_cache: dict = {}


def preprocess_u8(images_u8: torch.Tensor, channels_first: bool = False) -> torch.Tensor:
    """SegGptImageProcessor.preprocess for `images` / `prompt_images` (HF:image_processing_seggpt.py:134-252): uint8
    [n,c,c,3] (or [n,3,c,c]) square crops -> bicubic-antialias resize to 448 -> (x - 255 mean)/(255 std), float32
    [n,3,448,448]."""
    _need_cuda(images_u8)
    if images_u8.dtype != torch.uint8 or images_u8.ndim != 4:
        raise ValueError("preprocess_u8 expects a uint8 [n,c,c,3] or [n,3,c,c] tensor")
    n = images_u8.shape[0]
    crop = images_u8.shape[2]
    other = images_u8.shape[3] if channels_first else images_u8.shape[1]
    if other != crop:
        raise ValueError("preprocess_u8 handles square crops (the reference's crops are squares, src/util/ml_util.py:20-66)")
    dev = images_u8.device
    key = (crop, str(dev))
    if key not in _tv_cache:
        b, c, p = tv_bicubic_aa_table(crop, 448)
        _tv_cache[key] = (torch.from_numpy(b).to(dev), torch.from_numpy(c).to(dev), c.shape[1], p)
    bounds, coef, ksize, prec = _tv_cache[key]
    m255, s255 = _hf_mean_std_255()
    src = images_u8.contiguous()
    out = torch.empty((n, 3, 448, 448), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_preprocess_u8(_lib.ptr(src), 1 if channels_first else 0, n, crop, _lib.ptr(coef),
                                                 _lib.ptr(bounds), ksize, prec, _lib.f3(m255), _lib.f3(s255),
                                                 _lib.ptr(out), _lib.stream_ptr()), "bseg_preprocess_u8")
    return out


def colorize_resize_norm255(mask: torch.Tensor, palette_u8: torch.Tensor, out_size: int = 448) -> torch.Tensor:
    """SegGptImageProcessor.preprocess for segmentation-map `prompt_masks` (HF:image_processing_seggpt.py:100-131,
    175-215): class ids uint8 [B,c,c] -> palette colour -> NEAREST resize -> (rgb - 255 mean)/(255 std)."""
    _need_cuda(mask)
    if mask.ndim == 4:
        mask = mask.squeeze(1)
    B, H, W = mask.shape
    if H != W:
        raise ValueError("square masks only")
    dev = mask.device
    m8 = mask.to(torch.uint8).contiguous()
    pal = palette_u8.to(device=dev, dtype=torch.uint8).contiguous()
    idx = (_device_array("nearest_exact", (H, out_size), dev, lambda: torch_nearest_exact_index(H, out_size))
           if H != out_size else None)
    m255, s255 = _hf_mean_std_255()
    out = torch.empty((B, 3, out_size, out_size), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_colorize_resize_norm255(_lib.ptr(m8), _lib.ptr(pal), pal.shape[0], _lib.f3(m255),
                                                           _lib.f3(s255), _lib.ptr(idx), _lib.ptr(out), B, H, out_size,
                                                           _lib.stream_ptr()), "bseg_colorize_resize_norm255")
    return out


def postprocess_semantic(pred_masks: torch.Tensor, palette255: torch.Tensor, out_size: Optional[int] = None,
                         nodata: Optional[torch.Tensor] = None, dtype=torch.int64) -> torch.Tensor:
    """SegGptImageProcessor.post_process_semantic_segmentation (HF:image_processing_seggpt.py:254-321) for a batch,
    optionally with the nodata zeroing of src/predict_no_prompt.py:303.  pred_masks float32 [B,3,2H,W];
    palette255 float32 [C,3] (0..255).  Returns [B,out,out]."""
    _need_cuda(pred_masks)
    B, _, H2, W = pred_masks.shape
    H = H2 // 2
    out_size = H if out_size is None else int(out_size)
    dev = pred_masks.device
    idx = (_device_array("nearest", (H, out_size), dev, lambda: torch_nearest_index(H, out_size))
           if (out_size != H or out_size != W) else None)
    o8 = torch.empty((B, out_size, out_size), dtype=torch.uint8, device=dev) if dtype == torch.uint8 else None
    o64 = torch.empty((B, out_size, out_size), dtype=torch.int64, device=dev) if dtype == torch.int64 else None
    nd = nodata.to(torch.uint8).contiguous() if nodata is not None else None
    pred_c = pred_masks.contiguous()
    pal_c = palette255.to(device=dev, dtype=torch.float32).contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().bseg_postprocess_semantic(
            _lib.ptr(pred_c), _lib.ptr(pal_c), pal_c.shape[0], _lib.f3(IMAGE_MEAN), _lib.f3(IMAGE_STD), _lib.ptr(o8),
            _lib.ptr(o64), _lib.ptr(nd), _lib.ptr(idx), B, H, W, out_size, _lib.stream_ptr()),
            "bseg_postprocess_semantic")
    return o8 if o8 is not None else o64
