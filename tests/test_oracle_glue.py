"""CPU: pins oracle/glue_ref.py against fixtures produced by the REAL reference functions (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import glue_ref


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "glue_golden.npz")


def test_tif_image(g):
    out = glue_ref.tif_image_4band(g["tif_data"].copy(), g["tif_nodata"].copy())
    assert out.dtype == np.uint8 and np.array_equal(out, g["tif_out"])


def test_crop_tif(g):
    for i, b in enumerate(g["crop_boxes"]):
        ci, cn, _ = glue_ref.crop_tif(tuple(int(v) for v in b), g["tif_out"], g["tif_nodata"], None, 16)
        assert np.array_equal(ci, g[f"crop_img_{i}"])
        assert np.array_equal(cn, g[f"crop_nodata_{i}"])
    # fully outside: image all zero, nodata all one (padding value 1)
    assert g["crop_img_3"].max() == 0 and g["crop_nodata_3"].all()


def test_palettes(g):
    assert np.array_equal(np.array(glue_ref.build_palette(3)), g["build_palette_3"])
    assert np.array_equal(np.array(glue_ref.build_palette(7)), g["build_palette_7"])
    torch.manual_seed(42)
    pal = glue_ref.generate_random_rgb_palette(4, 3)
    assert np.array_equal(pal.numpy(), g["random_palette_seed42"])
    out = glue_ref.torch_apply_mask_rgb(pal, torch.from_numpy(g["apply_mask_in"]))
    assert np.array_equal(out.numpy(), g["apply_mask_out"])


@pytest.mark.parametrize("B", [1, 3])
def test_loss_as_written(g, B):
    pred, lab, yes = (torch.from_numpy(g[f"loss_{k}_B{B}"]) for k in ("pred", "labels", "yes"))
    out = glue_ref.seggpt_loss(pred, lab, yes, beta=0.01, per_sample=False)
    assert np.allclose(out.numpy(), g[f"loss_out_B{B}"], rtol=1e-6, atol=0)
    per = glue_ref.seggpt_loss(pred, lab, yes, beta=0.01, per_sample=True)
    if B == 1:
        assert torch.allclose(per, out)
    else:  # the broadcasting quirk of src/model.py:61 makes the two differ at B > 1
        assert abs(per.item() - out.item()) > 1e-3


def test_process_pred_masks(g):
    out = glue_ref.process_pred_masks(torch.from_numpy(g["decode_pred"]), torch.from_numpy(g["decode_palette_norm"]))
    assert np.array_equal(out.numpy(), g["decode_out"])


def test_accumulator(g):
    acc = glue_ref.AccumulatorRef((30, 44))
    oks = []
    for b, c in zip(g["vote_boxes"], g["vote_cls"]):
        oks.append(acc.update(tuple(int(v) for v in b), np.eye(4, dtype=np.uint8)[c]))
    assert oks == [True, True, True, True, False, True]  # box 4 is outside the scene ("Invalid crop!")
    assert np.array_equal(acc.counter, g["vote_counter"])
    assert np.array_equal(acc.argmax().astype(np.uint8), g["vote_argmax"])


def test_pil_bicubic_restatement(g):
    """The numpy restatement of PIL's 8-bit resampler (whose tables the CUDA ingest kernel consumes) is bit exact
    against PIL itself, down- and up-scaling."""
    assert np.array_equal(glue_ref.pil_bicubic_resize_ref(g["pil_in_64"], 56), g["pil_out_64_to_56"])
    assert np.array_equal(glue_ref.pil_bicubic_resize_ref(g["pil_in_64"], 100), g["pil_out_64_to_100"])
    from PIL import Image

    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(512, 512, 3)).astype(np.uint8)
    want = np.array(Image.fromarray(img).resize((448, 448), resample=Image.Resampling.BICUBIC))
    assert np.array_equal(glue_ref.pil_bicubic_resize_ref(img, 448), want)


def test_host_tables_match_oracle():
    """Host logic of the product (beach_seg_b200.ops) builds the same tables as the oracle restatement."""
    from beach_seg_b200 import ops

    for size in (112, 336, 512, 1024):
        b, k = ops.pil_bicubic_table(size, 448)
        rb, rk = glue_ref.pil_bicubic_coeffs(size, 448)
        assert np.array_equal(b, rb) and np.array_equal(k, rk)
    b, k = ops.pil_bicubic_table(448, 448)
    assert (b[:, 1] == 1).all() and (k == 1 << 22).all()
    import cv2

    for src, dst in ((448, 512), (448, 112), (448, 1024), (448, 336), (448, 500)):
        ramp = np.tile(np.arange(src, dtype=np.int32)[None, :], (src, 1))
        want = cv2.resize(ramp, (dst, dst), interpolation=cv2.INTER_NEAREST)[0]
        assert np.array_equal(ops.cv2_nearest_index(src, dst), want)
        assert np.array_equal(glue_ref.cv2_nearest_index(src, dst), want)


def test_merge_mosaic_hand_case():
    """merge_tifs accumulation (src/util/geo_util.py:410-419) on a case small enough to check by hand.
    rasterio is not in this image, so the reprojection step cannot be run; the accumulation is plain numpy."""
    data = np.zeros((2, 4, 1, 3), dtype=np.float32)
    data[0, :, 0, :] = [[10, 20, 30]] * 4
    data[1, :, 0, :] = [[30, 99, -7]] * 4
    yes = np.array([[[255, 255, 0]], [[255, 0, 0]]], dtype=np.uint8)
    mean, mask = glue_ref.merge_mosaic(data, yes)
    assert mean.dtype == np.float32 and mean.shape == (4, 1, 3)
    assert np.array_equal(mean[0, 0], np.array([20, 20, 0], dtype=np.float32))
    assert np.array_equal(mask, np.array([[False, False, True]]))


def _pil_overlay(img, pred, classes):
    """overlay_prediction exactly as the reference writes it (src/util/img_util.py:98-116), on the real Pillow."""
    from PIL import Image, ImageColor

    colors = {"nodata": None, "water": "yellow", "veg": "blue", "sand": "hotpink"}
    h, w, _ = img.shape
    layer = np.zeros((h, w, 4), dtype=np.uint8)
    for idx, name in enumerate(colors[c] for c in classes):
        if name is None:
            continue
        layer[pred == idx] = (*ImageColor.getrgb(name), int(255 * 0.3))
    blended = Image.alpha_composite(Image.fromarray(img).convert("RGBA"), Image.fromarray(layer, mode="RGBA"))
    return np.array(blended.convert("RGB"))


def test_overlay_prediction_matches_pillow():
    """Every (base value, class) pair: 256 grey levels x 4 classes, plus a random RGB image."""
    classes = ("nodata", "sand", "water", "veg")
    rng = np.random.default_rng(5)
    grey = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 4, axis=0).repeat(3, axis=2)
    pred = np.repeat(np.arange(4, dtype=np.uint8)[:, None], 256, axis=1)
    assert np.array_equal(glue_ref.overlay_prediction(grey, pred, classes), _pil_overlay(grey, pred, classes))
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    pred = rng.integers(0, 4, (97, 131), dtype=np.uint8)
    assert np.array_equal(glue_ref.overlay_prediction(img, pred, classes), _pil_overlay(img, pred, classes))
