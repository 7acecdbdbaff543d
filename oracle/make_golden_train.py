"""Generates tests/golden/train_golden.npz and tests/golden/ref_prompt_batch.pt.gz from the REAL reference code.
TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):   python -m oracle.make_golden_train

  train_golden.npz        learning-rate trajectories of the reference's own `PromptModel.configure_optimizers`
                          (src/model.py:385-428: AdamW, optional linear warm-up, CosineAnnealingLR stepped per epoch)
                          for three configurations; pins beach_seg_b200.model.PromptModel.configure_optimizers.
  ref_prompt_batch.pt.gz  `prompt_batch.pt` exactly as src/train.py:71-77 writes it: the reference's own
                          `PromptModel.create_trainable_params` (src/model.py:115-130) on two dataset items in the
                          layout of `BeachSegDataset.get_crop` (src/data.py:118-124), then the reference's own
                          `handle_item` (src/train.py:20-24) and `torch.save`.  gzip is applied afterwards, by this
                          script, only to keep the fixture small (the prompt images are piecewise constant);
                          pins beach_seg_b200.train.load_prompt_batch.
"""
from __future__ import annotations

import gzip
import importlib
import io
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle._ref_import import import_reference  # noqa: E402

LR_CASES = {
    # name: BeachSegConfig overrides
    "default": dict(epochs=6),
    "warmup": dict(epochs=10, warmup_epochs=3, lr=2e-3, init_lr=1e-4, min_lr=2e-5),
    "scaled": dict(epochs=8, warmup_epochs=2, batch_size=4, world_size=2, grad_accum_steps=2, base_lr_batch_size=4,
                   lr=1e-3, init_lr=5e-4, min_lr=1e-5),
}
LR_EPOCHS = 14  # past T_max on purpose: CosineAnnealingLR keeps oscillating, the mirror must do the same


def lr_trajectory(configure_optimizers, conf, n_epochs: int) -> np.ndarray:
    """lr at the start of epoch 0..n_epochs-1 when the scheduler is stepped once per epoch (Lightning, interval
    "epoch", frequency 1)."""
    p = torch.nn.Parameter(torch.zeros(3))
    fake = types.SimpleNamespace(conf=conf, parameters=lambda: [p])
    cfg = configure_optimizers(fake)
    opt, sched = cfg["optimizer"], cfg["lr_scheduler"]["scheduler"]
    assert cfg["lr_scheduler"]["interval"] == "epoch" and cfg["lr_scheduler"]["frequency"] == 1
    out = []
    for _ in range(n_epochs):
        out.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    return np.array(out, dtype=np.float64)


def prompt_items(n: int = 2):
    """Dataset items in the layout of BeachSegDataset.get_crop (src/data.py:118-124); piecewise-constant content so
    that the saved file compresses."""
    rng = np.random.default_rng(77)
    items = []
    for i in range(n):
        blocks = rng.integers(0, 256, size=(3, 14, 14)).astype(np.float32) / 255.0
        img = np.repeat(np.repeat(blocks, 32, axis=1), 32, axis=2)
        cls = np.repeat(np.repeat(rng.integers(0, 4, size=(14, 14)).astype(np.uint8), 32, axis=0), 32, axis=1)
        nodata = np.zeros((448, 448), dtype=bool)
        nodata[:32 * (i + 1), :64] = True
        items.append({"crop_idx": 3 + 4 * i, "date": f"2024010{i + 1}", "image": img.copy(), "mask": cls,
                      "nodata": nodata})
    return items


def main():
    gdir = ROOT / "tests" / "golden"
    ref = import_reference()
    ref_train = importlib.import_module("src.train")  # handle_item; its other imports are stubbed
    out = {}
    for name, kw in LR_CASES.items():
        conf = ref.config.BeachSegConfig(**kw)
        out[f"lr_{name}"] = lr_trajectory(ref.model.PromptModel.configure_optimizers, conf, LR_EPOCHS)
        print(name, out[f"lr_{name}"])
    np.savez_compressed(gdir / "train_golden.npz", **out)

    fake = types.SimpleNamespace()
    dm = types.SimpleNamespace(prompt_imgs=prompt_items())
    ref.model.PromptModel.create_trainable_params(fake, dm)           # src/model.py:115-130
    prompt_batch = {k: ref_train.handle_item(v) for k, v in fake.prompt_batch.items()}   # src/train.py:76
    buf = io.BytesIO()
    torch.save(prompt_batch, buf)                                      # src/train.py:77
    with gzip.GzipFile(gdir / "ref_prompt_batch.pt.gz", "wb", compresslevel=9, mtime=0) as f:
        f.write(buf.getvalue())
    print("prompt_batch keys:", {k: type(v).__name__ for k, v in prompt_batch.items()}, "raw bytes", buf.tell(),
          "gz bytes", (gdir / "ref_prompt_batch.pt.gz").stat().st_size)


if __name__ == "__main__":
    main()
