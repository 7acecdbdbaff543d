"""GPU parity of the train-time augmentation kernels (bseg_train_aug_fwd / _bwd through the C ABI) against the CPU
oracle (oracle/aug_ref.py: kornia's published ops restated; PARITY UNPINNED against kornia itself, see its header).
Bar: forward max |diff| < 2e-5 in normalised units (fp32, same op order; measured 0 on the CPU emulation of the same
code), masks bit-exact, gradient rel-L2 < 1e-3 apart from <= 0.5 % of elements on sub-gradient conventions."""
import pytest
import torch

from beach_seg_b200 import _lib, augment
from oracle import aug_ref
from tests._aug_common import busy_conf, compare_grad, draw, to_ref_params

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True)
def _needs_cuda(dev):
    """Skip (instead of failing with 'Found no NVIDIA driver') when a plain `pytest tests` runs on a host without a GPU."""


def _gpu_vs_oracle(conf, B, H, W, seed):
    aug, d, image, mask, noise = draw(conf, B, H, W, seed)
    img_dev = image.to(DEV).requires_grad_(True)
    out, out_mask = aug.apply(img_dev, mask.to(DEV), d, noise=noise.to(DEV))
    img_ref = image.clone().requires_grad_(True)
    ref, ref_mask = aug_ref.train_aug(img_ref, mask, to_ref_params(d, noise, conf))
    d_out = torch.randn(ref.shape, generator=torch.Generator().manual_seed(seed + 99))
    (g_ref,) = torch.autograd.grad(ref, img_ref, d_out)
    out.backward(d_out.to(DEV))
    torch.cuda.synchronize()
    return out.detach().cpu(), out_mask.cpu(), ref.detach(), ref_mask, img_dev.grad.cpu(), g_ref, d


@pytest.mark.parametrize("seed", range(4))
def test_train_aug_matches_oracle_at_model_size(seed):
    conf = busy_conf()
    out, out_mask, ref, ref_mask, g, g_ref, d = _gpu_vs_oracle(conf, 8, 448, 448, seed)
    assert torch.equal(out_mask, ref_mask)
    err = (out - ref).abs().max().item()
    exact = (out == ref).float().mean().item()
    print(f"seed {seed} order {d['order']}: max |diff| {err:.2e}, bit-identical {exact:.4%}")
    assert err < 2e-5
    frac, rel = compare_grad(g, g_ref, f"seed {seed}")
    print(f"  gradient: {frac:.4%} convention pixels, rel-L2 {rel:.2e}")


def test_train_aug_ragged_shapes_and_reference_defaults():
    """Non-square, odd sizes; the reference's default probabilities (most samples take no optional op)."""
    from beach_seg_b200.config import BeachSegConfig

    for (B, H, W, seed) in ((1, 7, 5, 0), (3, 33, 129, 1), (5, 64, 64, 2)):
        out, out_mask, ref, ref_mask, g, g_ref, _ = _gpu_vs_oracle(BeachSegConfig(), B, H, W, seed)
        assert torch.equal(out_mask, ref_mask)
        assert (out - ref).abs().max().item() < 2e-5
        compare_grad(g, g_ref, f"{B}x{H}x{W}", max_bad_frac=2e-2)


def test_train_aug_extremes():
    conf = busy_conf(sharpness=2.5, sharpness_p=1.0, brightness=0.6, contrast=0.8, saturation=1.0, hue=0.5)
    out, out_mask, ref, ref_mask, g, g_ref, _ = _gpu_vs_oracle(conf, 6, 96, 80, 7)
    assert torch.equal(out_mask, ref_mask)
    assert (out - ref).abs().max().item() < 2e-5
    compare_grad(g, g_ref, "extreme", max_bad_frac=2e-2)


def test_dict_call_like_the_reference_and_empty_batch():
    conf = busy_conf()
    aug = augment.TrainAug(conf, generator=torch.Generator().manual_seed(3))
    batch = {"image": torch.rand((4, 3, 64, 64), device=DEV), "mask": torch.randint(0, 4, (4, 1, 64, 64), device=DEV),
             "crop_idx": [1, 2, 3, 4]}
    out = aug(batch)
    assert out["image"].shape == (4, 3, 64, 64) and out["mask"].shape == (4, 1, 64, 64)
    assert out["mask"].dtype == batch["mask"].dtype and out["crop_idx"] == [1, 2, 3, 4]
    assert torch.isfinite(out["image"]).all()
    # determinism: the same draw twice gives the same bits
    a, _ = aug.apply(batch["image"], None, aug.last_params, noise=torch.zeros_like(batch["image"]))
    b, _ = aug.apply(batch["image"], None, aug.last_params, noise=torch.zeros_like(batch["image"]))
    assert torch.equal(a, b)
    empty = aug({"image": torch.empty((0, 3, 16, 16), device=DEV), "mask": torch.empty((0, 16, 16), dtype=torch.uint8,
                                                                                          device=DEV)})
    assert empty["image"].shape == (0, 3, 16, 16)
    with pytest.raises(_lib.BsegError):
        aug.apply(torch.rand((1, 3, 8, 8)), None, aug.sample_params(1, 8, 8))  # CPU tensor: no fallback


def test_prompt_gradient_flows_through_train_aug():
    """The reference's chain prompt parameter -> stack -> train_aug -> SegGPT -> pred_masks (src/model.py:194-207,
    245-251): the gradient reaching the prompt parameters through bseg_train_aug_bwd equals the gradient w.r.t. the
    augmented prompt (same CUDA model backward, leaf tensor) pushed through torch autograd of the oracle chain for the
    same parameter draw.  Small 5-layer backbone."""
    from beach_seg_b200 import synth
    from beach_seg_b200.model import PromptModel
    from beach_seg_b200.seggpt import SegGptB200
    from oracle.seggpt_ref import make_reference_model

    hf = make_reference_model(seed=0, stress=True, num_layers=5, merge_index=1, intermediate=(1, 2, 3, 4))
    backbone = SegGptB200.from_hf(hf, device=DEV)
    conf = busy_conf(gauss_p=0.0)  # the noise field is drawn on the device inside TrainAug.apply
    pm = PromptModel(conf, device=DEV, model=backbone)
    prompt_img01 = synth.smooth_image(2, seed=60)
    prompt_cls = synth.blocky_mask(2, seed=61)

    class DM:
        prompt_imgs = [{"image": prompt_img01[i], "mask": prompt_cls[i][None], "crop_idx": i} for i in range(2)]

    augs = augment.Augmentations(conf, generator=torch.Generator().manual_seed(11))
    DM.train_aug, DM.aug = augs.train_aug, augs.aug
    pm.post_init(DM)
    pm.create_trainable_params(DM)
    idx = [1, 0, 1]
    px = synth.normalize(synth.smooth_image(3, seed=62)).to(DEV)
    pal, _ = pm.create_palette(3, train=False)
    prompt_batch, prompt_masks = pm.prepare_prompt(idx, pal, train=True)
    d = augs.train_aug.last_params
    out = backbone(pixel_values=px, prompt_pixel_values=prompt_batch["image"], prompt_masks=prompt_masks,
                   embedding_type="instance")
    d_pred = torch.zeros_like(out.pred_masks)
    d_pred[:, :, 448:] = torch.randn((3, 3, 448, 448), generator=torch.Generator().manual_seed(5)).to(DEV)
    out.pred_masks.backward(d_pred)
    g_dev = [p.grad.detach().cpu() for p in pm.prompt_params_list]

    leaf = prompt_batch["image"].detach().clone().requires_grad_(True)
    out2 = backbone(pixel_values=px, prompt_pixel_values=leaf, prompt_masks=prompt_masks, embedding_type="instance")
    out2.pred_masks.backward(d_pred)
    torch.cuda.synchronize()
    params_cpu = [p.detach().cpu().clone().requires_grad_(True) for p in pm.prompt_params_list]
    stack = torch.stack([params_cpu[i] for i in idx])
    masks = torch.stack([prompt_cls[i] for i in idx])
    ref_aug, ref_mask = aug_ref.train_aug(stack, masks, to_ref_params(d, torch.zeros_like(stack), conf))
    assert (ref_aug.detach() - prompt_batch["image"].detach().cpu()).abs().max().item() < 2e-5
    assert torch.equal(ref_mask, prompt_batch["mask"].reshape(3, 448, 448).cpu())
    g_ref = torch.autograd.grad(ref_aug, params_cpu, leaf.grad.cpu())
    for i in range(2):
        compare_grad(g_dev[i], g_ref[i], f"prompt {i}")


def test_ingest_raw_output_feeds_the_training_batch_augmentation():
    """The training-batch path of the reference (src/data.py:93-96 dataset item in [0,1] -> train_aug in
    on_after_batch_transfer, src/data.py:295-313): `ingest_tiles(normalize=False)` stops after /255, and
    Normalize((x)) of it is the normalised ingest output bit for bit."""
    import numpy as np

    from beach_seg_b200 import ops, synth

    crop, Hs, Ws = 512, 700, 900
    scene = synth.scene_u16(Hs, Ws, seed=3)
    nodata = synth.nodata_wedge(Hs, Ws)
    sc = torch.from_numpy(scene.view(np.int16)).to(DEV)
    nd = torch.from_numpy(nodata).to(DEV)
    stats = ops.scene_stats(sc, nd)
    boxes = torch.tensor([[0, 0, crop, crop], [300, 150, 300 + crop, 150 + crop]], dtype=torch.int32, device=DEV)
    raw = ops.ingest_tiles(sc, nd, stats, boxes, crop, normalize=False)["image"]
    norm = ops.ingest_tiles(sc, nd, stats, boxes, crop)["image"]
    # compared on the CPU: torch's CUDA `x / 255` multiplies by a reciprocal, the kernel (like numpy in the reference)
    # divides
    raw_c, norm_c = raw.cpu(), norm.cpu()
    assert raw_c.min() >= 0 and raw_c.max() <= 1 and torch.equal((raw_c * 255).round() / 255, raw_c)   # u8 / 255 exactly
    mean = torch.tensor(ops.IMAGE_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(ops.IMAGE_STD).view(1, 3, 1, 1)
    assert torch.equal((raw_c - mean) / std, norm_c)
    # the reference's default pipeline on this batch: finite, mask moved with the image
    from beach_seg_b200.config import BeachSegConfig

    aug = augment.TrainAug(BeachSegConfig(), generator=torch.Generator().manual_seed(0))
    labels = synth.blocky_mask(2, seed=9).to(DEV)
    out = aug({"image": raw, "mask": labels})
    d = aug.last_params
    ref, ref_mask = aug_ref.train_aug(raw.cpu(), labels.cpu(), to_ref_params(d, torch.zeros(raw.shape), BeachSegConfig())) \
        if not bool(d["noise_apply"].any()) else (None, None)
    if ref is not None:
        assert (out["image"].cpu() - ref).abs().max().item() < 2e-5 and torch.equal(out["mask"].cpu(), ref_mask)


def test_train_aug_reproduces_the_committed_vector():
    """CUDA path vs tests/golden/aug_golden.npz (the restatement's output for a stored parameter draw)."""
    from tests.test_oracle_aug import _load_golden

    z, t, p = _load_golden()
    d = {"vflip": p.vflip, "hflip": p.hflip, "brightness": p.brightness, "contrast": p.contrast,
         "saturation": p.saturation, "hue": p.hue, "order": p.order, "sharp_apply": p.sharp_apply,
         "sharp_factor": p.sharp_factor, "erase_apply": p.erase_apply, "erase_box": p.erase_box,
         "noise_apply": p.noise_apply}
    from beach_seg_b200.config import BeachSegConfig

    conf = BeachSegConfig(gauss_mean=p.noise_mean, gauss_std=p.noise_std)
    aug = augment.TrainAug(conf)
    img = t("image").to(DEV).requires_grad_(True)
    out, out_mask = aug.apply(img, t("mask").to(DEV), d, noise=t("noise").to(DEV))
    out.backward(t("d_out").to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(out_mask.cpu(), t("out_mask"))
    assert (out.detach().cpu() - t("out")).abs().max().item() < 2e-5
    compare_grad(img.grad.cpu(), t("grad"), "golden")
