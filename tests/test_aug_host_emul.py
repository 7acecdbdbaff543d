"""The per-pixel functions of beach_seg_b200/csrc/augment_math.cuh (the code the CUDA kernels of augment.cu call) run
on the CPU through tests/host_emul/aug_emul.cpp and are compared with the oracle: forward against
oracle/aug_ref.train_aug, backward against torch autograd through it.  This checks the LOGIC without a GPU; the GPU
parity tests proper are in tests/test_gpu_augment.py (through the C ABI of libbseg.so)."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest
import torch

from beach_seg_b200 import _lib, augment
from oracle import aug_ref
from tests._aug_common import busy_conf, compare_grad, draw, to_ref_params

HERE = Path(__file__).resolve().parent / "host_emul"


@pytest.fixture(scope="module")
def emul():
    so = HERE / "_aug_emul.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-Wno-unknown-pragmas", "-shared", "-fPIC", "-o",
                    str(so), str(HERE / "aug_emul.cpp")], check=True)
    L = C.CDLL(str(so))
    sig = _lib.SIGNATURES
    L.emul_train_aug_fwd.restype = C.c_int
    L.emul_train_aug_fwd.argtypes = sig["bseg_train_aug_fwd"][1][:-1]  # same signature minus the stream
    L.emul_train_aug_bwd.restype = C.c_int
    L.emul_train_aug_bwd.argtypes = sig["bseg_train_aug_bwd"][1][:-1]
    return L


def _run(emul, conf, B, H, W, seed, order=None):
    aug, d, image, mask, noise = draw(conf, B, H, W, seed)
    if order is not None:
        d["order"] = order
    params = augment.pack_params(B, vflip=d["vflip"], hflip=d["hflip"], brightness=d["brightness"],
                                 contrast=d["contrast"], saturation=d["saturation"], hue=d["hue"],
                                 sharp_apply=d["sharp_apply"], sharp_factor=d["sharp_factor"],
                                 erase_apply=d["erase_apply"], erase_box=d["erase_box"], noise_apply=d["noise_apply"])
    rc, out, out_mask, colour = augment._raw_fwd(emul.emul_train_aug_fwd, image, mask, params, d["order"], noise,
                                                 conf.gauss_mean, conf.gauss_std, aug.mean, aug.std, ())
    assert rc == 0
    img_ref = image.clone().requires_grad_(True)
    ref, ref_mask = aug_ref.train_aug(img_ref, mask, to_ref_params(d, noise, conf))
    d_out = torch.randn(ref.shape, generator=torch.Generator().manual_seed(seed + 99))
    (g_ref,) = torch.autograd.grad(ref, img_ref, d_out)
    rc, g = augment._raw_bwd(emul.emul_train_aug_bwd, image, params, d["order"], aug.std, colour, d_out, ())
    assert rc == 0
    return out, out_mask, ref.detach(), ref_mask, g, g_ref, d


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("H,W", [(40, 36), (23, 37)])  # W % 4 == 0: the four-pixel finish passes; otherwise the scalar ones
def test_forward_and_gradient_match_the_oracle(emul, seed, H, W):
    conf = busy_conf()
    out, out_mask, ref, ref_mask, g, g_ref, d = _run(emul, conf, 6, H, W, seed)
    assert torch.equal(out_mask, ref_mask)
    err = (out - ref).abs().max().item()
    assert err < 2e-5, (err, d["order"])   # normalised units (1/std ~ 4.4x the [0,1] image scale)
    compare_grad(g, g_ref, f"seed {seed} order {d['order']}")


def test_every_colour_order(emul):
    import itertools

    conf = busy_conf(sharpness_p=0.0, erasing_p=0.0, gauss_p=0.0)
    for i, order in enumerate(itertools.permutations(range(4))):
        out, _, ref, _, g, g_ref, _ = _run(emul, conf, 2, 24, 24, 100 + i, order=order)
        assert (out - ref).abs().max().item() < 2e-5, order
        compare_grad(g, g_ref, f"order {order}")


def test_sharpness_outside_unit_interval_and_extreme_colours(emul):
    conf = busy_conf(sharpness=2.5, sharpness_p=1.0, brightness=0.6, contrast=0.8, saturation=1.0, hue=0.5)
    out, out_mask, ref, ref_mask, g, g_ref, _ = _run(emul, conf, 6, 32, 32, 7)
    assert torch.equal(out_mask, ref_mask)
    assert (out - ref).abs().max().item() < 2e-5
    compare_grad(g, g_ref, "extreme", max_bad_frac=2e-2)


def test_bad_order_is_rejected(emul):
    conf = busy_conf()
    aug, d, image, mask, noise = draw(conf, 1, 8, 8, 0)
    params = augment.pack_params(1)
    rc, *_ = augment._raw_fwd(emul.emul_train_aug_fwd, image, mask, params, (0, 1, 1, 3), noise, 0.0, 0.1, aug.mean,
                              aug.std, ())
    assert rc != 0


def test_emulated_kernel_code_reproduces_the_committed_vector(emul):
    import numpy as np

    from tests.test_oracle_aug import _load_golden

    z, t, p = _load_golden()
    B = z["image"].shape[0]
    params = augment.pack_params(B, vflip=p.vflip, hflip=p.hflip, brightness=p.brightness, contrast=p.contrast,
                                 saturation=p.saturation, hue=p.hue, sharp_apply=p.sharp_apply,
                                 sharp_factor=p.sharp_factor, erase_apply=p.erase_apply, erase_box=p.erase_box,
                                 noise_apply=p.noise_apply)
    rc, out, out_mask, colour = augment._raw_fwd(emul.emul_train_aug_fwd, t("image"), t("mask"), params, p.order,
                                                 t("noise"), p.noise_mean, p.noise_std, aug_ref.IMAGE_MEAN,
                                                 aug_ref.IMAGE_STD, ())
    assert rc == 0 and torch.equal(out_mask, t("out_mask"))
    assert (out - t("out")).abs().max().item() < 2e-5
    rc, g = augment._raw_bwd(emul.emul_train_aug_bwd, t("image"), params, p.order, aug_ref.IMAGE_STD, colour, t("d_out"), ())
    assert rc == 0
    compare_grad(g, t("grad"), "golden")
