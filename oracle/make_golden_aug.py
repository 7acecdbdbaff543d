"""Generates tests/golden/aug_golden.npz.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_aug

Unlike glue_golden.npz / seggpt_golden.npz this fixture is NOT produced by reference code: kornia, which holds the
arithmetic of the reference's train-time augmentation (src/data.py:195-224), is neither under /root/reference nor in
this image, so parity of oracle/aug_ref.py with kornia stays UNPINNED (see its header).  The fixture is the
restatement's own output for one fully specified parameter draw (every optional op on, every per-sample value stored
in the file), so that (i) an accidental change of the restatement is caught, (ii) the CUDA path can be compared with a
committed vector (tests/test_gpu_augment.py), and (iii) a maintainer who HAS kornia can replay the stored parameters
through `K.AugmentationSequential(...)(..., params=...)` and pin the restatement in one step.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import aug_ref  # noqa: E402


def fixed_draw(B=3, H=24, W=20):
    g = torch.Generator().manual_seed(2024)
    image = (torch.rand((B, 3, H, W), generator=g) * 255).round() / 255
    mask = torch.randint(0, 4, (B, H, W), generator=g, dtype=torch.uint8)
    noise = torch.randn((B, 3, H, W), generator=g)
    p = aug_ref.AugParams(
        vflip=torch.tensor([True, False, True]), hflip=torch.tensor([False, True, True]),
        brightness=torch.tensor([1.07, 0.93, 1.0]), contrast=torch.tensor([0.95, 1.08, 1.02]),
        saturation=torch.tensor([1.1, 0.9, 1.05]), hue=torch.tensor([0.08, -0.06, 0.1]), order=(2, 0, 3, 1),
        sharp_apply=torch.tensor([True, True, False]), sharp_factor=torch.tensor([0.35, 0.8, 0.5]),
        erase_apply=torch.tensor([True, False, True]), erase_box=torch.tensor([[3, 5, 6, 4], [0, 0, 1, 1], [10, 2, 5, 9]]),
        erase_value=0.0, noise_apply=torch.tensor([False, True, True]), noise=noise, noise_mean=0.0, noise_std=0.1)
    return image, mask, p


def main():
    image, mask, p = fixed_draw()
    img = image.clone().requires_grad_(True)
    out, out_mask = aug_ref.train_aug(img, mask, p)
    d_out = torch.randn(out.shape, generator=torch.Generator().manual_seed(7))
    (grad,) = torch.autograd.grad(out, img, d_out)
    path = ROOT / "tests" / "golden" / "aug_golden.npz"
    np.savez_compressed(
        path, image=image.numpy(), mask=mask.numpy(), noise=p.noise.numpy(), d_out=d_out.numpy(),
        out=out.detach().numpy(), out_mask=out_mask.numpy(), grad=grad.numpy(),
        vflip=p.vflip.numpy(), hflip=p.hflip.numpy(), brightness=p.brightness.numpy(), contrast=p.contrast.numpy(),
        saturation=p.saturation.numpy(), hue=p.hue.numpy(), order=np.array(p.order), sharp_apply=p.sharp_apply.numpy(),
        sharp_factor=p.sharp_factor.numpy(), erase_apply=p.erase_apply.numpy(), erase_box=p.erase_box.numpy(),
        noise_apply=p.noise_apply.numpy(), noise_mean=np.float32(p.noise_mean), noise_std=np.float32(p.noise_std))
    print(f"wrote {path} ({path.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
