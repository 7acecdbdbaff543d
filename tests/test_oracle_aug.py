"""The kornia restatement (oracle/aug_ref.py) cannot be pinned against kornia itself (absent from the image and from
/root/reference: PARITY UNPINNED); these tests pin what can be pinned: algebraic properties and the one op that has an
independent implementation in the image (torchvision's adjust_sharpness == kornia.enhance.sharpness by construction)."""
import math

import torch

from oracle import aug_ref
from tests._aug_common import busy_conf, draw, to_ref_params


def _identity_params(B, H, W):
    z = torch.zeros(B, dtype=torch.bool)
    one = torch.ones(B)
    return aug_ref.AugParams(vflip=z, hflip=z, brightness=one, contrast=one, saturation=one, hue=torch.zeros(B),
                             order=(0, 1, 2, 3), sharp_apply=z, sharp_factor=one, erase_apply=z,
                             erase_box=torch.zeros((B, 4), dtype=torch.int64), erase_value=0.0, noise_apply=z,
                             noise=torch.zeros((B, 3, H, W)), noise_mean=0.0, noise_std=0.1)


def test_identity_parameters_reduce_to_normalize():
    img = torch.rand((2, 3, 16, 12), generator=torch.Generator().manual_seed(0))
    mask = torch.randint(0, 4, (2, 16, 12), dtype=torch.uint8)
    out, m = aug_ref.train_aug(img, mask, _identity_params(2, 16, 12))
    mean = torch.tensor(aug_ref.IMAGE_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(aug_ref.IMAGE_STD).view(1, 3, 1, 1)
    assert torch.allclose(out, (img - mean) / std, atol=1e-5)  # the HSV round trip of saturation=1 / hue=0 is not exact
    assert torch.equal(m, mask)


def test_hsv_round_trip_and_ranges():
    img = torch.rand((3, 3, 20, 20), generator=torch.Generator().manual_seed(1))
    hsv = aug_ref.rgb_to_hsv(img)
    assert hsv[:, 0].min() >= 0 and hsv[:, 0].max() < 2 * math.pi + 1e-6
    assert hsv[:, 1].min() >= 0 and hsv[:, 1].max() <= 1
    assert torch.allclose(aug_ref.hsv_to_rgb(hsv), img, atol=1e-6)
    # pure colours land on their sector
    prim = torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]]).view(3, 3, 1, 1)
    assert torch.allclose(aug_ref.rgb_to_hsv(prim)[:, 0, 0, 0], torch.tensor([0.0, 2 * math.pi / 3, 4 * math.pi / 3]), atol=1e-6)


def test_hue_full_turn_is_identity_and_saturation_zero_is_grey():
    img = torch.rand((2, 3, 10, 10), generator=torch.Generator().manual_seed(2))
    assert torch.allclose(aug_ref.adjust_hue(img, torch.full((2,), 2 * math.pi)), img, atol=2e-6)
    grey = aug_ref.adjust_saturation(img, torch.zeros(2))
    assert torch.allclose(grey, img.max(1, keepdim=True).values.expand_as(img), atol=1e-6)


def test_sharpness_matches_torchvision():
    from torchvision.transforms.v2 import functional as TF

    img = torch.rand((3, 3, 24, 17), generator=torch.Generator().manual_seed(3))
    for f in (0.0, 0.3, 1.0, 1.7):
        ours = aug_ref.sharpness(img, torch.full((3,), f))
        assert torch.allclose(ours, TF.adjust_sharpness(img, f), atol=1e-6), f


def test_flips_are_involutions_and_move_the_mask_with_the_image():
    conf = busy_conf(sharpness_p=0.0, erasing_p=0.0, gauss_p=0.0, hue=0.0, saturation=0.0, contrast=0.0, brightness=0.0)
    _, d, img, mask, noise = draw(conf, 6, 12, 10, seed=4)
    d["order"] = (0, 1, 3, 2)
    p = to_ref_params(d, noise, conf)
    out, m = aug_ref.train_aug(img, mask, p, mean=(0, 0, 0), std=(1, 1, 1))
    out2, m2 = aug_ref.train_aug(out, m, p, mean=(0, 0, 0), std=(1, 1, 1))
    assert torch.allclose(out2, img, atol=1e-5) and torch.equal(m2, mask)
    # the image of the mask, pushed through the same flips as data, equals the flipped mask
    as_img = mask.float()[:, None].expand(-1, 3, -1, -1) / 4
    out3, _ = aug_ref.train_aug(as_img, None, p, mean=(0, 0, 0), std=(1, 1, 1))
    assert torch.allclose(out3[:, 0] * 4, m.float(), atol=1e-4)


def test_erase_zeroes_image_and_mask_inside_the_box_only():
    conf = busy_conf(erasing_p=1.0, sharpness_p=0.0, gauss_p=0.0, vertical_flip=0.0, horizontal_flip=0.0)
    _, d, img, mask, noise = draw(conf, 4, 30, 30, seed=5)
    p = to_ref_params(d, noise, conf)
    out, m = aug_ref.train_aug(img, mask + 1, p, mean=(0, 0, 0), std=(1, 1, 1))
    for b in range(4):
        x, y, w, h = d["erase_box"][b].tolist()
        assert (out[b, :, y:y + h, x:x + w] == 0).all() and (m[b, y:y + h, x:x + w] == 0).all()
        assert int((m[b] == 0).sum()) == w * h


def test_sample_params_follow_the_documented_distributions():
    """Host logic of augment.TrainAug.sample_params: ranges, shapes, reproducibility, RandomErasing geometry."""
    from beach_seg_b200.augment import TrainAug, pack_params
    from beach_seg_b200.config import BeachSegConfig

    conf = BeachSegConfig()
    B, H, W = 4096, 448, 448
    d = TrainAug(conf, generator=torch.Generator().manual_seed(0)).sample_params(B, H, W)
    d2 = TrainAug(conf, generator=torch.Generator().manual_seed(0)).sample_params(B, H, W)
    for k in d:
        assert (d[k] == d2[k]) if isinstance(d[k], tuple) else torch.equal(d[k], d2[k]), k
    assert sorted(d["order"]) == [0, 1, 2, 3]
    for key, p in (("vflip", conf.vertical_flip), ("hflip", conf.horizontal_flip), ("sharp_apply", conf.sharpness_p),
                   ("erase_apply", conf.erasing_p), ("noise_apply", conf.gauss_p)):
        assert abs(d[key].float().mean().item() - p) < 0.03, key
    for key, c in (("brightness", conf.brightness), ("contrast", conf.contrast), ("saturation", conf.saturation)):
        assert d[key].min() >= 1 - c and d[key].max() <= 1 + c and abs(d[key].mean().item() - 1) < 0.01
    assert d["hue"].min() >= -conf.hue and d["hue"].max() <= conf.hue
    assert d["sharp_factor"].min() >= 0 and d["sharp_factor"].max() <= conf.sharpness
    x, y, w, h = d["erase_box"].unbind(1)
    assert (w >= 1).all() and (h >= 1).all() and (x >= 0).all() and (y >= 0).all()
    assert (x + w <= W).all() and (y + h <= H).all()
    frac = (w * h).float() / (H * W)
    assert frac.min() > 0.8 * conf.erasing_scale[0] and frac.max() < 1.2 * conf.erasing_scale[1]
    ratio = h.float() / w.float()
    assert ratio.min() > 0.25 and ratio.max() < 3.6
    P = pack_params(B, **{k: v for k, v in d.items() if k != "order"})
    assert P.shape == (B, 16) and P.dtype == torch.float32
    assert torch.equal(P[:, 2], d["brightness"] - 1) and torch.equal(P[:, 5], d["hue"] * 2 * math.pi)


def _load_golden():
    import numpy as np
    from pathlib import Path

    z = np.load(Path(__file__).resolve().parent / "golden" / "aug_golden.npz")
    t = lambda k: torch.from_numpy(z[k])
    p = aug_ref.AugParams(
        vflip=t("vflip"), hflip=t("hflip"), brightness=t("brightness"), contrast=t("contrast"),
        saturation=t("saturation"), hue=t("hue"), order=tuple(int(v) for v in z["order"]), sharp_apply=t("sharp_apply"),
        sharp_factor=t("sharp_factor"), erase_apply=t("erase_apply"), erase_box=t("erase_box"), erase_value=0.0,
        noise_apply=t("noise_apply"), noise=t("noise"), noise_mean=float(z["noise_mean"]), noise_std=float(z["noise_std"]))
    return z, t, p


def test_restatement_reproduces_the_committed_vector():
    """tests/golden/aug_golden.npz (oracle/make_golden_aug.py): the restatement's own output for a stored parameter
    draw -- a regression pin, NOT a kornia pin (kornia is not available here)."""
    z, t, p = _load_golden()
    img = t("image").clone().requires_grad_(True)
    out, out_mask = aug_ref.train_aug(img, t("mask"), p)
    (grad,) = torch.autograd.grad(out, img, t("d_out"))
    assert torch.allclose(out.detach(), t("out"), atol=1e-6) and torch.equal(out_mask, t("out_mask"))
    assert torch.allclose(grad, t("grad"), atol=1e-5)
