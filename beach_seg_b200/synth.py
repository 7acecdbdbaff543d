"""Seeded synthetic inputs of the shapes the reference path consumes (there is no network for real Planet scenes
or the pretrained checkpoint).  Used by bench.py, the tests and the golden-vector generator.  CPU / numpy only.
"""
from __future__ import annotations

import numpy as np
import torch

BAND_MEAN = (600.0, 900.0, 1100.0, 2500.0)   # Dove surface-reflectance-like
BAND_STD = (200.0, 300.0, 400.0, 800.0)
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)
PALETTE3 = ((0, 0, 0), (255, 255, 255), (255, 255, 127), (255, 127, 255))  # build_palette(3)


def scene_u16(height: int, width: int, seed: int = 7) -> np.ndarray:
    """uint16 [4, H, W] band-planar scene: clip(N(mu_b, sigma_b) * low-frequency field, 1, 10000)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.meshgrid(np.arange(height, dtype=np.float32), np.arange(width, dtype=np.float32), indexing="ij")
    field = 1.0 + 0.35 * np.sin(yy / 97.0 + 0.3) * np.cos(xx / 131.0 - 0.7)
    out = np.empty((4, height, width), dtype=np.uint16)
    for b in range(4):
        v = rng.normal(BAND_MEAN[b], BAND_STD[b], size=(height, width)).astype(np.float32) * field
        out[b] = np.clip(v, 1, 10000).astype(np.uint16)
    return out


def nodata_wedge(height: int, width: int, frac: float = 0.05) -> np.ndarray:
    """bool [H, W]: a wedge along the left/top border covering about `frac` of the scene."""
    yy, xx = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    return (xx / max(width, 1) + yy / max(height, 1)) < np.sqrt(2 * frac)


def smooth_image(batch: int, seed: int, size: int = 448) -> torch.Tensor:
    """float32 [B,3,size,size] in [0,1]: band-limited random field + fine noise (image-like statistics)."""
    g = torch.Generator().manual_seed(seed)
    low = torch.rand((batch, 3, 14, 14), generator=g)
    img = torch.nn.functional.interpolate(low, size=(size, size), mode="bilinear", align_corners=False)
    img = img + 0.08 * torch.randn((batch, 3, size, size), generator=g)
    return img.clamp_(0.0, 1.0)


def blocky_mask(batch: int, seed: int, num_classes: int = 4, size: int = 448) -> torch.Tensor:
    """uint8 [B,size,size]: randint(0, num_classes) on a (size/16)^2 grid, upsampled x16."""
    g = torch.Generator().manual_seed(seed)
    m = torch.randint(0, num_classes, (batch, size // 16, size // 16), generator=g, dtype=torch.uint8)
    return m.repeat_interleave(16, dim=1).repeat_interleave(16, dim=2).contiguous()


def normalize(x: torch.Tensor) -> torch.Tensor:
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32).view(1, 3, 1, 1)
    std = torch.tensor(IMAGE_STD, dtype=torch.float32).view(1, 3, 1, 1)
    return (x - mean) / std


def model_inputs(batch: int, seed: int = 123, size: int = 448):
    """(pixel_values, prompt_pixel_values, prompt_masks), each float32 [B,3,size,size], normalised like the reference
    does (images: /255 + ImageNet mean/std; prompt masks: build_palette(3) colours, same normalisation)."""
    px = normalize(smooth_image(batch, seed, size))
    ppx = normalize(smooth_image(batch, seed + 1, size))
    cls = blocky_mask(batch, seed + 2, size=size).long()
    pal = torch.tensor(PALETTE3, dtype=torch.float32) / 255.0
    pm = normalize(pal[cls].permute(0, 3, 1, 2).contiguous())
    return px, ppx, pm


def tile_boxes(n_tiles: int, crop: int, width: int) -> np.ndarray:
    """int32 [n,4] (xmin,ymin,xmax,ymax): a row-major grid of non-overlapping crop x crop tiles."""
    per_row = max(width // crop, 1)
    out = np.zeros((n_tiles, 4), dtype=np.int32)
    for i in range(n_tiles):
        x0, y0 = (i % per_row) * crop, (i // per_row) * crop
        out[i] = (x0, y0, x0 + crop, y0 + crop)
    return out


def sliding_boxes(height: int, width: int, crop: int, stride: int) -> np.ndarray:
    """Dense sliding window with overlap (BASELINE config 3: crop 512, stride 448 -> 18 x 9 on 8000 x 4000)."""
    xs = list(range(0, max(width - crop, 0) + stride, stride))
    ys = list(range(0, max(height - crop, 0) + stride, stride))
    return np.array([(x, y, x + crop, y + crop) for y in ys for x in xs], dtype=np.int32)
