"""CPU oracle for the tensor glue around the SegGPT call on the beach_seg hot path.  TEST INFRASTRUCTURE ONLY
(see oracle/seggpt_ref.py for who may import this).

`src.model`, `src.data`, `src.predict*`, `src.util.geo_util` of the reference cannot be imported as they are in
this image (lightning, kornia, rasterio, geopandas, shapely, omegaconf ... are absent), so the lines on the hot path
are restated here one for one, each function naming the reference lines it follows.  The restatement is pinned
against the REAL reference functions, imported with stubbed third-party modules by oracle/make_golden.py, through
the fixtures in tests/golden/ (tests/test_oracle_glue.py).

Allowed deviations: kornia `Normalize` -> two torch ops; Lightning `self.device` / `log_dict` dropped.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

IMAGE_MEAN = (0.485, 0.456, 0.406)  # HF:image_processing_seggpt.py:76
IMAGE_STD = (0.229, 0.224, 0.225)   # HF:image_processing_seggpt.py:77


# ------------------------------------------------------------------------------------------------------------
# ingest
# ------------------------------------------------------------------------------------------------------------
def tif_image_4band(data: np.ndarray, nodata: np.ndarray) -> np.ndarray:
    """src/util/geo_util.py:449-470, 4-band branch.  data: (4,H,W) float32, nodata: (H,W) bool -> (H,W,3) uint8."""
    img = np.zeros((3, *data.shape[1:]), dtype=data.dtype)
    img[0] = data[3]
    img[1] = data[2]
    img[2] = data[:2].mean(axis=0)
    min_val = img[:, ~nodata].min()
    img = img.clip(min_val, 3000 + min_val) - min_val
    img -= img[:, ~nodata].min()
    for i in range(3):
        img[i] /= img[i].max()
        img[i][nodata] = 0
    img = img.transpose((1, 2, 0)).copy()
    return np.array(img * 255, dtype=np.uint8)


def merge_mosaic(dst_data: np.ndarray, dst_yesdata: np.ndarray):
    """src/util/geo_util.py:410-419 — the accumulation of merge_tifs once every raster is on the output grid.
    dst_data: (N,C,H,W) float32; dst_yesdata: (N,H,W) uint8 (rasterio masks, 0/255) -> (mean (C,H,W) float32, nodata (H,W) bool)."""
    mask_f = dst_yesdata.astype(dst_data.dtype)
    w = mask_f[:, None, :, :]
    weighted_sum = (dst_data * w).sum(axis=0)
    weights = w.sum(axis=0)
    mean = np.divide(weighted_sum, weights, out=np.full_like(weighted_sum, 0.0), where=weights != 0)
    mask = ~np.any(dst_yesdata, axis=0)
    return mean, mask


CLASS_COLORS_RGB = {"nodata": None, "water": (255, 255, 0), "veg": (0, 0, 255), "sand": (255, 105, 180)}
"""src/util/img_util.py:12 with ImageColor.getrgb applied (yellow, blue, hotpink)."""


def overlay_prediction(img: np.ndarray, pred: np.ndarray, classes) -> np.ndarray:
    """src/util/img_util.py:98-116 in integer numpy.  Pillow 12.2 (libImaging/AlphaComposite.c) composites with
    7 precision bits; for an opaque destination: coef1 = a*128, coef2 = 255*128 - coef1,
    out = div255(src*coef1 + dst*coef2 + (0x80 << 7)) >> 7 with div255(t) = ((t >> 8) + t) >> 8; a == 0 copies dst.
    Pinned against PIL.Image.alpha_composite in tests/test_oracle_glue.py.  img (H,W,3) uint8, pred (H,W) -> (H,W,3)."""
    alpha = int(255 * 0.3)
    out = img.copy()
    for cls_idx, name in enumerate(classes):
        rgb = CLASS_COLORS_RGB[name]
        if rgb is None:
            continue
        m = pred == cls_idx
        dst = img[m].astype(np.uint32)
        coef1 = np.uint32(alpha << 7)
        coef2 = np.uint32((255 << 7) - (alpha << 7))
        t = np.array(rgb, dtype=np.uint32)[None, :] * coef1 + dst * coef2 + np.uint32(0x80 << 7)
        out[m] = ((((t >> 8) + t) >> 8) >> 7).astype(np.uint8)
    return out


def padded_crop(arr: np.ndarray, xmin: int, ymin: int, xmax: int, ymax: int, crop_size: int, value=0) -> np.ndarray:
    """src/util/geo_util.py:316-341."""
    if arr.ndim == 3:
        h, w, c = arr.shape
        padded = np.full((crop_size, crop_size, c), fill_value=value, dtype=arr.dtype)
    else:
        h, w = arr.shape
        padded = np.full((crop_size, crop_size), fill_value=value, dtype=arr.dtype)
    x0, x1 = max(xmin, 0), min(xmax, w)
    y0, y1 = max(ymin, 0), min(ymax, h)
    ystart = y0 - ymin
    yend = ystart + (y1 - y0)
    xstart = x0 - xmin
    xend = xstart + (x1 - x0)
    padded[ystart:yend, xstart:xend] = arr[y0:y1, x0:x1]
    return padded


def crop_tif(crop, img, nodata, label, crop_size):
    """src/util/geo_util.py:297-313."""
    xmin, ymin, xmax, ymax = crop
    crop_img = padded_crop(img, xmin, ymin, xmax, ymax, crop_size)
    crop_nodata = padded_crop(nodata, xmin, ymin, xmax, ymax, crop_size, value=1)
    crop_label = padded_crop(label, xmin, ymin, xmax, ymax, crop_size) if label is not None else None
    return crop_img, crop_nodata, crop_label


def get_crop_image(crop_img_u8: np.ndarray, inpt_size: int = 448) -> np.ndarray:
    """src/data.py:93-96,119: PIL BICUBIC resize (if sizes differ), /255, HWC->CHW float32."""
    from PIL import Image

    im = Image.fromarray(crop_img_u8)
    if inpt_size != crop_img_u8.shape[0]:
        im = im.resize((inpt_size, inpt_size), resample=Image.Resampling.BICUBIC)
    out = np.array(im).astype(np.float32) / 255.0
    return out.transpose((2, 0, 1)).copy()


def resize_nearest_pil(arr: np.ndarray, size: int) -> np.ndarray:
    """src/data.py:98-113: label / nodata resized with PIL NEAREST."""
    from PIL import Image

    if size == arr.shape[0]:
        return np.array(arr)
    return np.array(Image.fromarray(arr).resize((size, size), resample=Image.Resampling.NEAREST))


def normalize(x: torch.Tensor) -> torch.Tensor:
    """K.Normalize(mean, std) (src/data.py:226-229,342-343): (x - mean) / std per channel, float32."""
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGE_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def denormalize(x: torch.Tensor) -> torch.Tensor:
    """K.Denormalize (src/data.py:345-346): x * std + mean."""
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGE_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return x * std + mean


# --- PIL's 8-bit resampler restated (libImaging/Resample.c), used to pin the coefficient tables the CUDA ingest
# --- kernel consumes; get_crop_image() above calls PIL itself like the reference does.
def _bicubic_filter(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_bicubic_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc of PIL Resample.c -> (bounds [out,2], kk int32 [out,ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def pil_bicubic_resize_ref(img_u8: np.ndarray, out_size: int) -> np.ndarray:
    """Two-pass (horizontal, then vertical) 8-bit fixed-point resample of a square HWC uint8 image."""
    in_size = img_u8.shape[0]
    bounds, kk = pil_bicubic_coeffs(in_size, out_size)
    src = img_u8.astype(np.int64)
    tmp = np.zeros((in_size, out_size, img_u8.shape[2]), dtype=np.int64)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = (1 << 21) + np.tensordot(src[:, x0:x0 + n, :], kk[xx, :n].astype(np.int64), axes=([1], [0]))
        tmp[:, xx, :] = np.clip(acc >> 22, 0, 255)
    out = np.zeros((out_size, out_size, img_u8.shape[2]), dtype=np.int64)
    for yy in range(out_size):
        y0, n = bounds[yy]
        acc = (1 << 21) + np.tensordot(kk[yy, :n].astype(np.int64), tmp[y0:y0 + n], axes=([0], [0]))
        out[yy] = np.clip(acc >> 22, 0, 255)
    return out.astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------
# palettes / colourise
# ------------------------------------------------------------------------------------------------------------
def build_palette(num_labels: int):
    """src/util/ml_util.py:72-89."""
    base = int(num_labels ** (1 / 3)) + 1
    margin = 256 // base
    color_list = [(0, 0, 0)]
    for location in range(num_labels):
        num_seq_r = location // base**2
        num_seq_g = (location % base**2) // base
        num_seq_b = location % base
        color_list.append((255 - num_seq_r * margin, 255 - num_seq_g * margin, 255 - num_seq_b * margin))
    return color_list


def generate_random_rgb_palette(num_labels: int, batch_size: int, generator=None) -> torch.Tensor:
    """src/util/ml_util.py:99-111 (draws from the global torch RNG when generator is None, like the reference)."""
    lut = torch.randint(low=0, high=256, size=(batch_size, num_labels, 3), dtype=torch.uint8, generator=generator)
    lut[:, 0] = 0
    return lut


def torch_apply_mask_rgb(palette: torch.Tensor, inp: torch.Tensor) -> torch.Tensor:
    """src/util/ml_util.py:114-132."""
    if inp.ndim == 3:
        inp = inp.unsqueeze(1)
    mask = inp.squeeze(1).to(torch.long)
    B = mask.shape[0]
    rgb = palette[torch.arange(B)[:, None, None], mask]
    return rgb.permute(0, 3, 1, 2).to(dtype=torch.float32) / 255.0


def create_palette(num_classes: int, batch_size: int, train: bool, generator=None):
    """PromptModel.create_palette (src/model.py:215-231)."""
    if train:
        batch_palette = generate_random_rgb_palette(num_classes, batch_size, generator)
    else:
        palette = torch.Tensor(build_palette(num_classes - 1))
        batch_palette = torch.stack([palette for _ in range(batch_size)])
    palettes = []
    for pal in batch_palette:
        pn = normalize_palette_entry(pal, num_classes)
        palettes.append(pn)
    return batch_palette, torch.stack(palettes)


def normalize_palette_entry(pal: torch.Tensor, num_classes: int) -> torch.Tensor:
    """src/model.py:223-228: normalize(palette.view(C,3,1,1).float()/255) squeezed -> (C,3)."""
    x = pal.view((num_classes, 3, 1, 1)).to(torch.float32) / 255
    return normalize(x).squeeze(-1).squeeze(-1)


# ------------------------------------------------------------------------------------------------------------
# decode / loss
# ------------------------------------------------------------------------------------------------------------
def process_pred_masks(in_pred_masks: torch.Tensor, batch_palette_norm: torch.Tensor) -> torch.Tensor:
    """PromptModel.process_pred_masks (src/model.py:155-175)."""
    H = in_pred_masks.shape[2] // 2
    masks = in_pred_masks[:, :, H:, :]
    out = []
    for idx, mask in enumerate(masks):
        palette = batch_palette_norm[idx]
        channels, height, width = mask.shape
        dist = mask.permute(1, 2, 0).reshape(height, width, 1, channels)
        dist = dist - palette.view(1, 1, 4, 3)
        dist = torch.pow(dist, 2)
        dist = torch.sum(dist, dim=-1)
        out.append(dist.argmin(dim=-1))
    return torch.stack(out, dim=0)


def seggpt_loss(pred_masks: torch.Tensor, labels: torch.Tensor, yesdata: torch.Tensor, beta: float = 0.01,
                per_sample: bool = False) -> torch.Tensor:
    """SegGptLoss.forward (src/model.py:45-64).  per_sample=False reproduces the code AS WRITTEN, including the
    `keep_mask.unsqueeze(1)` broadcast to (B,B,C,2H,W); per_sample=True is the B=1-equivalent intended form."""
    B, C, H2, W = pred_masks.shape
    H = H2 // 2
    blank = torch.zeros((B, C, H, W), dtype=pred_masks.dtype)
    label_mask = torch.concat([blank, labels], dim=2)
    keep_mask = torch.concat([blank, yesdata.expand((-1, C, -1, -1)).to(pred_masks.dtype)], dim=2)
    loss = F.smooth_l1_loss(pred_masks, label_mask, reduction="none", beta=beta)
    if per_sample:
        loss = loss * keep_mask
    else:
        loss = loss * keep_mask.unsqueeze(1).to(loss.dtype)
    return loss.sum() / keep_mask.sum()


def predict_tail(pred_mask: np.ndarray, crop_size: int, num_classes: int = 4):
    """src/predict.py:258-260: cv2 INTER_NEAREST back to crop size, one-hot uint8."""
    import cv2

    pm = cv2.resize(pred_mask, (crop_size, crop_size), interpolation=cv2.INTER_NEAREST)
    return pm, np.eye(num_classes, dtype=np.uint8)[pm]


def cv2_nearest_index(src: int, dst: int) -> np.ndarray:
    """Source index per destination index of cv2.resize(INTER_NEAREST): min(floor(x * (1/(dst/src))), src-1)."""
    inv_scale = dst / src
    ifx = 1.0 / inv_scale
    return np.array([min(int(math.floor(x * ifx)), src - 1) for x in range(dst)], dtype=np.int32)


# ------------------------------------------------------------------------------------------------------------
# vote stitching
# ------------------------------------------------------------------------------------------------------------
class AccumulatorRef:
    """Tensor part of Accumulator (src/predict.py:55-159; src/predict_no_prompt.py:109-186)."""

    def __init__(self, out_shape, num_classes: int = 4):
        self.out_shape = out_shape
        self.counter = np.zeros((*out_shape, num_classes), dtype=np.uint8)
        self.img = np.zeros((*out_shape, 3), dtype=np.uint8)

    def update(self, crop, one_hot_pred: np.ndarray, img_crop: np.ndarray | None = None) -> bool:
        h, w = self.out_shape
        xmin, ymin, xmax, ymax = crop
        dy0, dy1 = max(ymin, 0), min(ymax, h)
        dx0, dx1 = max(xmin, 0), min(xmax, w)
        sy0 = dy0 - ymin
        sy1 = sy0 + (dy1 - dy0)
        sx0 = dx0 - xmin
        sx1 = sx0 + (dx1 - dx0)
        if sy1 <= sy0 or sx1 <= sx0:
            return False
        if img_crop is not None:
            self.img[dy0:dy1, dx0:dx1] = img_crop[sy0:sy1, sx0:sx1]
        self.counter[dy0:dy1, dx0:dx1] += one_hot_pred[sy0:sy1, sx0:sx1]
        return True

    def argmax(self) -> np.ndarray:
        return np.argmax(self.counter, axis=2)


def training_step_ref(hf_model, prompt_params, prompt_masks_cls, batch_image, batch_mask, generator, beta=0.01,
                      num_classes: int = 4, per_sample: bool = False):
    """PromptModel.training_step (src/model.py:233-269) with the infer-style augmentation (Normalize only; the kornia
    train augmentations are stochastic host ops outside the hot path): random palette from the GLOBAL torch RNG,
    label colourise + normalise, prompt choice from the module's private generator, prompt gather, HF forward, the
    reference's own SegGptLoss.  `prompt_params`: list of [3,448,448] tensors in [0,1] (requires_grad);
    `prompt_masks_cls`: list of [1,448,448] class-id tensors; batch_mask: [B,1,448,448].  Returns (loss, prompt_idx)."""
    B = batch_mask.shape[0]
    batch_palette, _ = create_palette(num_classes, B, train=True)                       # :235
    color_mask_norm = normalize(torch_apply_mask_rgb(batch_palette, batch_mask))        # :238-239
    prompt_idx = torch.randint(0, len(prompt_params), (B,), generator=generator)        # :242
    idx = prompt_idx.flatten().tolist()
    prompt_img = normalize(torch.stack([prompt_params[i] for i in idx], dim=0))         # :194 + aug (Normalize)
    prompt_mask = torch.stack([prompt_masks_cls[i] for i in idx], dim=0)
    prompt_color = normalize(torch_apply_mask_rgb(batch_palette, prompt_mask))          # :210-211
    out = hf_model(pixel_values=batch_image, labels=color_mask_norm, prompt_pixel_values=prompt_img,
                   prompt_masks=prompt_color, embedding_type="instance")                # :245-251
    loss = seggpt_loss(out.pred_masks, color_mask_norm, batch_mask != 0, beta, per_sample=per_sample)  # :255
    return loss, prompt_idx
