"""Mirror of the reference's `BeachSegConfig` (src/config.py:15-78): same field names and defaults, because the
hot-path code reads them (`classes`, `seed`, `crop_size`, `inpt_size`, `loss_beta`, `checkpoint`, `lr`, ...).
`resample` is kept as the PIL enum value name to avoid a hard PIL dependency on the device path."""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path

CLASSES = ("nodata", "sand", "water", "veg")  # src/config.py:7-12


@dataclass
class BeachSegConfig:
    project: str = "beach_seg"
    seed: int = 42
    # the reference's defaults are the author's private directories (src/config.py:19-20); neutral ones here
    data: Path = Path("data")
    model_training_root: Path = Path("results")
    classes: tuple = CLASSES
    devices: tuple = ("auto",)
    accelerator: str = "auto"
    deterministic: bool = False
    num_viz_images: int = 9
    viz_size: int = 224

    epochs: int = 1
    debug: bool = False
    world_size: int = 1
    grad_accum_steps: int = 1
    log_every_n_steps: int = 10
    precision: str = "32-true"
    workers: int = -1
    batch_size: int = 1

    checkpoint: str = "BAAI/seggpt-vit-large"

    monitor_metric: str = "val/f1"
    monitor_mode: str = "max"

    crop_size: int = 112
    inpt_size: int = 448
    resample: str = "BICUBIC"

    horizontal_flip: float = 0.5
    vertical_flip: float = 0.5
    hue: float = 0.1
    saturation: float = 0.1
    contrast: float = 0.1
    brightness: float = 0.1
    scale: tuple = (0.4, 1.0)
    sharpness: float = 1.0
    sharpness_p: float = 0.2
    erasing_scale: tuple = (0.02, 0.05)
    erasing_p: float = 0.1
    gauss_mean: float = 0.0
    gauss_std: float = 0.1
    gauss_p: float = 0.1
    channel_shift_limit: float = 0.01
    channel_shift_p: float = 0.2
    mosaic_p: float = 0.0
    jigsaw_grid: tuple = (2, 2)
    jigsaw_p: float = 0.0

    lr: float = 1e-3
    loss_beta: float = 0.01
    base_lr_batch_size: int = 1
    warmup_epochs: int = 0
    init_lr: float = 5e-04
    min_lr: float = 5e-04
    optimizer: str = "adamw"
    scheduler: str = "cosine"
    ema_alpha = 0.99
