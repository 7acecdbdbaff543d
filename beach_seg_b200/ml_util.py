"""Drop-ins for the tensor helpers of src/util/ml_util.py (same names, argument meaning and error behaviour)."""
from __future__ import annotations

import torch

from . import _lib, ops
from .seggpt import SegGptB200


def load_model(checkpoint: str, device: str | torch.device = "cuda:0", image_size: int = 448, **kw) -> SegGptB200:
    """src/util/ml_util.py:7-13: `from_pretrained(checkpoint)`, freeze, eval.  The reference then wraps the module in
    torch.compile; here the frozen backbone is packed once into the kernel layouts of libbseg.so instead.
    `checkpoint="random-init:<seed>"` builds HF's seeded random init (the only option without network access);
    with `image_size=512` that is the native-resolution variant `SegGptConfig(image_size=(1024, 512))`."""
    from transformers import SegGptConfig, SegGptForImageSegmentation

    if checkpoint.startswith("random-init"):
        seed = int(checkpoint.split(":")[1]) if ":" in checkpoint else 0
        torch.manual_seed(seed)
        cfg = SegGptConfig() if image_size == 448 else SegGptConfig(image_size=[2 * image_size, image_size])
        hf = SegGptForImageSegmentation(cfg)
    else:
        hf = SegGptForImageSegmentation.from_pretrained(checkpoint)
    for p in hf.parameters():  # freeze the backbone
        p.requires_grad_(False)
    model = SegGptB200.from_hf(hf.eval(), device=device, **kw)
    del hf
    return model.eval()


def build_palette(num_labels: int) -> list[tuple[int, int, int]]:
    """src/util/ml_util.py:72-89 (host logic)."""
    base = int(num_labels ** (1 / 3)) + 1
    margin = 256 // base
    color_list = [(0, 0, 0)]
    for location in range(num_labels):
        num_seq_r = location // base**2
        num_seq_g = (location % base**2) // base
        num_seq_b = location % base
        color_list.append((255 - num_seq_r * margin, 255 - num_seq_g * margin, 255 - num_seq_b * margin))
    return color_list


def generate_random_rgb_palette(num_labels: int, batch_size: int, device) -> torch.Tensor:
    """src/util/ml_util.py:99-111.  The reference runs on CPU and therefore consumes the GLOBAL CPU generator; to stay
    drop-in (same palette for the same torch.manual_seed) the draw is always made on the CPU and then moved."""
    lut = torch.randint(low=0, high=256, size=(batch_size, num_labels, 3), dtype=torch.uint8)
    lut[:, 0] = 0
    return lut.to(device)


def torch_apply_mask_rgb(palette: torch.Tensor, input: torch.Tensor) -> torch.Tensor:
    """src/util/ml_util.py:114-132: class ids -> palette colour / 255, float32 (B,3,H,W) (colourise kernel with
    mean 0 / std 1, which is exactly `rgb / 255`)."""
    if input.ndim == 3:
        input = input.unsqueeze(1)
    mask = input.squeeze(1)
    B, H, W = mask.shape
    m8 = mask.to(torch.uint8).contiguous()
    pal = palette.to(device=mask.device, dtype=torch.uint8).contiguous()
    out = torch.empty((B, 3, H, W), dtype=torch.float32, device=mask.device)
    if not mask.is_cuda:
        raise _lib.BsegError("torch_apply_mask_rgb: CUDA tensors only (no CPU fallback)")
    with torch.cuda.device(mask.device):
        _lib.check(_lib.lib().bseg_colorize_norm(_lib.ptr(m8), _lib.ptr(pal), pal.shape[1], _lib.f3((0, 0, 0)),
                                                 _lib.f3((1, 1, 1)), _lib.ptr(out), B, H, W, _lib.stream_ptr()),
                   "bseg_colorize_norm")
    return out
