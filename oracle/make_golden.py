"""Generates tests/golden/*.npz from the REAL reference code.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

  glue_golden.npz    outputs of the reference's own functions (imported through oracle/_ref_import.py) on seeded
                     inputs: tif_image, padded_crop/crop_tif, build_palette, generate_random_rgb_palette,
                     torch_apply_mask_rgb, SegGptLoss, PromptModel.process_pred_masks, Accumulator.update/argmax,
                     PIL BICUBIC resize (what BeachSegDataset.get_crop calls).
  seggpt_golden.npz  a strided slice + statistics of pred_masks of the HF SegGPT module (the un-vendored dependency
                     that holds the arithmetic, transformers 5.5.0) for the seeded synthetic inputs of
                     beach_seg_b200/synth.py, default init (seed 0) and stress init.

The fixtures pin oracle/glue_ref.py and oracle/seggpt_ref.py (tests/test_oracle_*.py, CPU) and are also compared
directly with the CUDA path (tests/test_gpu_*.py).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle._ref_import import import_reference  # noqa: E402
from beach_seg_b200 import synth  # noqa: E402  (seeded synthetic inputs only; no kernels involved)


def glue_golden(ref) -> dict:
    out = {}
    rng = np.random.default_rng(11)
    # ---- tif_image (src/util/geo_util.py:449-470) ----
    data = synth.scene_u16(40, 56, seed=5).astype(np.float32)
    nodata = synth.nodata_wedge(40, 56)
    out["tif_data"] = data
    out["tif_nodata"] = nodata
    out["tif_out"] = ref.geo_util.tif_image(data.copy(), nodata.copy())
    # ---- crop_tif / padded_crop (src/util/geo_util.py:297-341) ----
    boxes = np.array([[4, 6, 20, 22], [-5, -3, 11, 13], [48, 30, 64, 46], [100, 100, 116, 116]], dtype=np.int32)
    out["crop_boxes"] = boxes
    for i, b in enumerate(boxes):
        ci, cn, _ = ref.geo_util.crop_tif(tuple(int(v) for v in b), out["tif_out"], nodata, None, 16)
        out[f"crop_img_{i}"] = ci
        out[f"crop_nodata_{i}"] = cn
    # ---- palettes (src/util/ml_util.py:72-132) ----
    out["build_palette_3"] = np.array(ref.ml_util.build_palette(3), dtype=np.int64)
    out["build_palette_7"] = np.array(ref.ml_util.build_palette(7), dtype=np.int64)
    torch.manual_seed(42)
    pal = ref.ml_util.generate_random_rgb_palette(4, 3, "cpu")
    out["random_palette_seed42"] = pal.numpy()
    mask = torch.from_numpy(rng.integers(0, 4, size=(3, 1, 12, 10)).astype(np.uint8))
    out["apply_mask_in"] = mask.numpy()
    out["apply_mask_out"] = ref.ml_util.torch_apply_mask_rgb(pal, mask).numpy()
    # ---- SegGptLoss (src/model.py:40-64) ----
    for B in (1, 3):
        pred = torch.from_numpy(rng.normal(0, 1.0, size=(B, 3, 16, 8)).astype(np.float32))
        pred[:, :, 8:, :2] *= 0.005  # exercise the |d| < beta branch
        labels = torch.from_numpy(rng.normal(0, 1.0, size=(B, 3, 8, 8)).astype(np.float32))
        labels[:, :, :, :2] *= 0.005
        yes = torch.from_numpy(rng.random((B, 1, 8, 8)) > 0.3)
        out[f"loss_pred_B{B}"] = pred.numpy()
        out[f"loss_labels_B{B}"] = labels.numpy()
        out[f"loss_yes_B{B}"] = yes.numpy()
        out[f"loss_out_B{B}"] = ref.model.SegGptLoss(0.01)(pred, labels, yes).numpy()
    # ---- PromptModel.process_pred_masks (src/model.py:155-175) ----
    fake_self = types.SimpleNamespace(device="cpu")
    pm = torch.from_numpy(rng.normal(0, 1.5, size=(2, 3, 16, 8)).astype(np.float32))
    pal_norm = torch.from_numpy(rng.normal(0, 1.5, size=(2, 4, 3)).astype(np.float32))
    pal_norm[1, 2] = pal_norm[1, 1]  # a tie: first minimum must win
    out["decode_pred"] = pm.numpy()
    out["decode_palette_norm"] = pal_norm.numpy()
    out["decode_out"] = ref.model.PromptModel.process_pred_masks(fake_self, pm, pal_norm).numpy()
    # ---- Accumulator.update + argmax (src/predict.py:55-159,100) ----
    with tempfile.TemporaryDirectory() as td:
        acc = ref.predict.Accumulator((30, 44), Path(td), None, None, ("nodata", "sand", "water", "veg"))
        vote_boxes = np.array([[0, 0, 16, 16], [8, 8, 24, 24], [-6, -4, 10, 12], [36, 20, 52, 36], [60, 60, 76, 76],
                               [8, 8, 24, 24]], dtype=np.int32)
        cls = rng.integers(0, 4, size=(len(vote_boxes), 16, 16)).astype(np.int64)
        for b, c in zip(vote_boxes, cls):
            one_hot = np.eye(4, dtype=np.uint8)[c]
            acc.update("d0", tuple(int(v) for v in b), one_hot, np.zeros((16, 16, 3), np.uint8), None)
        out["vote_boxes"] = vote_boxes
        out["vote_cls"] = cls.astype(np.uint8)
        out["vote_counter"] = acc.current_pred_counter.copy()
        out["vote_argmax"] = np.argmax(acc.current_pred_counter, axis=2).astype(np.uint8)
        acc.current_pred_counter = None  # skip save_current (file output, out of scope) in __exit__-less use
    # ---- PIL BICUBIC resize as called by BeachSegDataset.get_crop (src/data.py:93-96) ----
    from PIL import Image

    small = rng.integers(0, 256, size=(64, 64, 3)).astype(np.uint8)
    out["pil_in_64"] = small
    out["pil_out_64_to_56"] = np.array(Image.fromarray(small).resize((56, 56), resample=Image.Resampling.BICUBIC))
    out["pil_out_64_to_100"] = np.array(Image.fromarray(small).resize((100, 100), resample=Image.Resampling.BICUBIC))
    return out


def seggpt_golden() -> dict:
    from oracle.seggpt_ref import make_reference_model

    out = {}
    for tag, stress in (("default", False), ("stress", True)):
        model = make_reference_model(seed=0, stress=stress)
        px, ppx, pm = synth.model_inputs(batch=1, seed=123)
        with torch.no_grad():
            pred = model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance").pred_masks
        out[f"{tag}_slice"] = pred[:, :, ::16, ::16].numpy().copy()
        out[f"{tag}_row500"] = pred[0, :, 500, :].numpy().copy()
        out[f"{tag}_mean_abs"] = np.array([pred.abs().mean().item()], dtype=np.float64)
        print(tag, "pred mean|x| =", pred.abs().mean().item(), "std =", pred.std().item())
    return out


def main():
    gdir = ROOT / "tests" / "golden"
    gdir.mkdir(parents=True, exist_ok=True)
    ref = import_reference()
    print("reference imported; stubbed third-party modules:", ref.stubbed)
    np.savez_compressed(gdir / "glue_golden.npz", **glue_golden(ref))
    print("wrote", gdir / "glue_golden.npz", os.path.getsize(gdir / "glue_golden.npz"), "bytes")
    if "--skip-model" not in sys.argv:
        np.savez_compressed(gdir / "seggpt_golden.npz", **seggpt_golden())
        print("wrote", gdir / "seggpt_golden.npz", os.path.getsize(gdir / "seggpt_golden.npz"), "bytes")


if __name__ == "__main__":
    main()
