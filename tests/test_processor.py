"""The HF image-processor path of src/predict_no_prompt.py.

CPU: the restated coefficient table of torchvision's uint8 bicubic-antialias resize and the nearest-index table are
checked bit for bit against torchvision / torch themselves.  GPU: the device preprocessing / post-processing against
the real `transformers.SegGptImageProcessor` (the un-vendored dependency the reference calls), bit-exact, and the
whole no-prompt tile body against the reference pipeline."""
import numpy as np
import pytest
import torch

from beach_seg_b200 import ops, synth


def _two_pass(img, out, bounds, coef, prec):
    n = img.shape[0]
    src = img.astype(np.int64)
    tmp = np.zeros((n, out, 3), np.int64)
    for x in range(out):
        x0, c = bounds[x]
        acc = (1 << (prec - 1)) + np.tensordot(src[:, x0:x0 + c, :], coef[x, :c].astype(np.int64), axes=([1], [0]))
        tmp[:, x, :] = np.clip(acc >> prec, 0, 255)
    o = np.zeros((out, out, 3), np.int64)
    for y in range(out):
        y0, c = bounds[y]
        acc = (1 << (prec - 1)) + np.tensordot(coef[y, :c].astype(np.int64), tmp[y0:y0 + c], axes=([0], [0]))
        o[y] = np.clip(acc >> prec, 0, 255)
    return o.astype(np.uint8)


@pytest.mark.parametrize("n", [336, 512, 1024, 100])
def test_tv_bicubic_aa_table_bit_exact_vs_torchvision(n):
    import torchvision.transforms.v2.functional as tvF
    from torchvision.transforms import InterpolationMode

    img = np.random.default_rng(n).integers(0, 256, (n, n, 3), dtype=np.uint8)
    t = torch.from_numpy(img).permute(2, 0, 1).contiguous()
    want = tvF.resize(t, (448, 448), interpolation=InterpolationMode.BICUBIC, antialias=True).permute(1, 2, 0).numpy()
    bounds, coef, prec = ops.tv_bicubic_aa_table(n, 448)
    assert np.array_equal(_two_pass(img, 448, bounds, coef, prec), want)


def test_torch_nearest_index_matches_torch():
    for s, d in ((448, 336), (448, 1024), (336, 448), (1024, 448), (448, 512), (112, 448)):
        x = torch.arange(s, dtype=torch.float32).view(1, 1, s, 1).expand(1, 1, s, s).contiguous()
        y = torch.nn.functional.interpolate(x, size=(d, d), mode="nearest")[0, 0, :, 0].long().numpy()
        assert np.array_equal(y, ops.torch_nearest_index(s, d)), (s, d)
        y = torch.nn.functional.interpolate(x, size=(d, d), mode="nearest-exact")[0, 0, :, 0].long().numpy()
        assert np.array_equal(y, ops.torch_nearest_exact_index(s, d)), (s, d)


def _hf_processor():
    from transformers import SegGptImageProcessor

    return SegGptImageProcessor()


@pytest.mark.gpu
@pytest.mark.parametrize("crop", [336, 1024])
def test_preprocess_matches_hf_processor(dev, crop):
    from beach_seg_b200.processor import load_processor

    rng = np.random.default_rng(crop)
    imgs = [rng.integers(0, 256, (crop, crop, 3), dtype=np.uint8) for _ in range(2)]
    masks = [rng.integers(0, 4, (crop, crop), dtype=np.uint8) for _ in range(2)]
    hf = _hf_processor()
    want = hf.preprocess(images=imgs, prompt_images=imgs[::-1], prompt_masks=masks, num_labels=3, return_tensors="pt",
                         data_format="channels_first")
    ours = load_processor("BAAI/seggpt-vit-large", device=dev)
    got = ours.preprocess(images=imgs, prompt_images=imgs[::-1], prompt_masks=masks, num_labels=3, return_tensors="pt",
                          data_format="channels_first")
    for k in ("pixel_values", "prompt_pixel_values", "prompt_masks"):
        assert got[k].shape == want[k].shape and got[k].dtype == want[k].dtype
        assert torch.equal(got[k].cpu(), want[k]), k  # bit-exact: integer resampling + the same two float32 ops
    assert ours.image_mean == list(hf.image_mean) and ours.image_std == list(hf.image_std)


@pytest.mark.gpu
@pytest.mark.parametrize("target", [336, 448, 1024])
def test_post_process_matches_hf_processor(dev, target):
    from beach_seg_b200.processor import SegGptOutputLike, load_processor

    g = torch.Generator().manual_seed(target)
    pred = torch.randn((1, 3, 896, 448), generator=g) * 1.5
    hf = _hf_processor()
    want = hf.post_process_semantic_segmentation(SegGptOutputLike(pred), [(target, target)], num_labels=3)[0]
    ours = load_processor(device=dev)
    got = ours.post_process_semantic_segmentation(SegGptOutputLike(pred.to(dev)), [(target, target)], num_labels=3)[0]
    assert got.shape == want.shape and got.dtype == want.dtype
    assert torch.equal(got.cpu(), want)


@pytest.mark.gpu
def test_no_prompt_tile_body_matches_reference_pipeline(dev):
    """src/predict_no_prompt.py:283-304 for two tiles x two prompts in ONE launch (5-layer stress model) against the
    reference pipeline (real HF processor + HF model, one tile at a time)."""
    from beach_seg_b200.predict import NoPromptPredictor
    from beach_seg_b200.processor import load_processor
    from beach_seg_b200.seggpt import SegGptB200
    from oracle.seggpt_ref import make_reference_model

    small = dict(num_layers=5, merge_index=1, intermediate=(1, 2, 3, 4))
    hf_model = make_reference_model(seed=1, stress=True, **small)
    model = SegGptB200.from_hf(hf_model, device=dev)
    crop = 336
    rng = np.random.default_rng(5)
    to_u8 = lambda t: (t.permute(0, 2, 3, 1) * 255).to(torch.uint8).numpy()
    tiles = to_u8(torch.nn.functional.interpolate(synth.smooth_image(2, 70), size=(crop, crop)))
    prompts = to_u8(torch.nn.functional.interpolate(synth.smooth_image(2, 71), size=(crop, crop)))
    pmasks = [np.ascontiguousarray(synth.blocky_mask(2, 72)[i, :crop, :crop].numpy()) for i in range(2)]
    nodata = rng.random((2, crop, crop)) < 0.05
    hf = _hf_processor()
    want = []
    pin = [hf.preprocess(prompt_images=[prompts[i]], prompt_masks=[pmasks[i]], num_labels=3, return_tensors="pt",
                         data_format="channels_first") for i in range(2)]
    with torch.no_grad():
        for t in range(2):
            inp = hf.preprocess(images=[tiles[t]] * 2, num_labels=3, return_tensors="pt", data_format="channels_first")
            out = hf_model(pixel_values=inp["pixel_values"],
                           prompt_pixel_values=torch.concat([p["prompt_pixel_values"] for p in pin]),
                           prompt_masks=torch.concat([p["prompt_masks"] for p in pin]), embedding_type="instance",
                           feature_ensemble=True)
            out.pred_masks = out.pred_masks.mean(dim=0).unsqueeze(0)
            pred = hf.post_process_semantic_segmentation(out, [(crop, crop)], num_labels=3)[0].numpy()
            pred[nodata[t]] = 0
            want.append(pred)
    want = np.stack(want)
    ours = load_processor(device=dev)
    pin_d = [ours.preprocess(prompt_images=[prompts[i]], prompt_masks=[pmasks[i]], num_labels=3) for i in range(2)]
    ppx = torch.concat([p["prompt_pixel_values"] for p in pin_d] * 2)
    pm = torch.concat([p["prompt_masks"] for p in pin_d] * 2)
    got = NoPromptPredictor(model, ours, crop).predict_tiles(torch.from_numpy(tiles).to(dev),
                                                             torch.from_numpy(nodata).to(dev), ppx, pm).cpu().numpy()
    flips = float((got != want).mean())
    print(f"[no-prompt tile body] class-map flips vs reference pipeline: {flips * 100:.3f} %")
    assert got.shape == (2, crop, crop) and np.all(got[nodata] == 0)
    assert flips < 0.01  # bf16 backbone: only pixels on a palette decision boundary may flip
    fast = NoPromptPredictor(model, ours, crop, query_half_only=True).predict_tiles(
        torch.from_numpy(tiles).to(dev), torch.from_numpy(nodata).to(dev), ppx, pm).cpu().numpy()
    assert np.array_equal(fast, got)  # the decoder's prompt half is never read by the post-processing


@pytest.mark.gpu
def test_no_prompt_full_size_batch_properties(dev):
    """BASELINE config 5 at full size: 16 tiles of 1024x1024 x 2 prompts (32 model samples, 24-layer backbone) in one
    launch.  Size-independent properties: every tile's class map is bit-identical to running that tile alone (the
    ensemble couples only the prompts of one tile), nodata pixels are 0, classes stay in range."""
    from beach_seg_b200.ml_util import load_model
    from beach_seg_b200.predict import NoPromptPredictor
    from beach_seg_b200.processor import load_processor

    n, P, crop = 16, 2, 1024
    model = load_model("random-init:0", device=dev)
    proc = load_processor(device=dev)
    rng = np.random.default_rng(9)
    to_u8 = lambda t: (t.permute(0, 2, 3, 1) * 255).to(torch.uint8)
    tiles = to_u8(torch.nn.functional.interpolate(synth.smooth_image(n, 80), size=(crop, crop), mode="bilinear")).to(dev)
    nodata = torch.from_numpy(rng.random((n, crop, crop)) < 0.03).to(dev)
    prompts = to_u8(torch.nn.functional.interpolate(synth.smooth_image(P, 81), size=(crop, crop))).numpy()
    pmasks = [np.ascontiguousarray(np.kron(synth.blocky_mask(P, 82)[i, :256, :256].numpy(), np.ones((4, 4), np.uint8)))
              for i in range(P)]
    pin = [proc.preprocess(prompt_images=[prompts[i]], prompt_masks=[pmasks[i]], num_labels=3) for i in range(P)]
    ppx1 = torch.concat([p["prompt_pixel_values"] for p in pin])
    pm1 = torch.concat([p["prompt_masks"] for p in pin])
    pred = NoPromptPredictor(model, proc, crop)
    got = pred.predict_tiles(tiles, nodata, ppx1.repeat(n, 1, 1, 1), pm1.repeat(n, 1, 1, 1))
    assert got.shape == (n, crop, crop) and got.dtype == torch.uint8
    assert int(got.max()) <= 3 and not got[nodata].any()
    for i in (0, 7, 15):
        one = pred.predict_tiles(tiles[i:i + 1], nodata[i:i + 1], ppx1, pm1)
        assert torch.equal(one[0], got[i])
