"""GPU: the bandwidth-bound glue kernels (through the C ABI wrappers in beach_seg_b200.ops) against the oracle
restatement of the reference lines and the committed golden fixtures.  Integer / index work is bit exact."""
import numpy as np
import pytest
import torch

from beach_seg_b200 import _lib, ops, synth
from oracle import glue_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(golden_dir / "glue_golden.npz")


def test_layernorm(dev):
    gen = torch.Generator().manual_seed(1)
    x = torch.randn((1000, 1024), generator=gen) * 3 + 0.5
    w = 1 + 0.1 * torch.randn(1024, generator=gen)
    b = 0.1 * torch.randn(1024, generator=gen)
    want = torch.nn.functional.layer_norm(x, (1024,), w, b, 1e-6)
    xd, wd, bd = x.to(dev), w.to(dev), b.to(dev)
    out = torch.zeros((1000, 4096), dtype=torch.bfloat16, device=dev)
    view = out[:, 1024:2048]
    _lib.check(_lib.lib().bseg_layernorm1024(_lib.ptr(xd), 1024, _lib.ptr(wd), _lib.ptr(bd), _lib.ptr(view), 4096,
                                             1000, 1e-6, _lib.stream_ptr()))
    torch.cuda.synchronize()
    got = out[:, 1024:2048].float().cpu()
    assert (got - want).abs().max().item() < 2e-2  # bf16 output rounding of |x| <~ 4
    assert torch.equal(got, want.to(torch.bfloat16).float()) or (got - want.to(torch.bfloat16).float()).abs().max() < 0.04
    assert out[:, :1024].abs().max().item() == 0 and out[:, 2048:].abs().max().item() == 0


def test_colorize_norm_bit_exact(dev, g):
    pal = torch.from_numpy(g["random_palette_seed42"])
    mask = torch.from_numpy(g["apply_mask_in"])
    want = glue_ref.normalize(torch.from_numpy(g["apply_mask_out"]))  # golden from the real torch_apply_mask_rgb
    got = ops.colorize_norm(mask.to(dev), pal.to(dev)).cpu()
    assert torch.equal(got, want)
    # full-size seeded case against the oracle restatement
    m = synth.blocky_mask(3, seed=4)
    torch.manual_seed(7)
    pal = glue_ref.generate_random_rgb_palette(4, 3)
    want = glue_ref.normalize(glue_ref.torch_apply_mask_rgb(pal, m))
    assert torch.equal(ops.colorize_norm(m.to(dev), pal.to(dev)).cpu(), want)


def test_decode_palette_bit_exact(dev, g):
    pred = torch.from_numpy(g["decode_pred"])
    paln = torch.from_numpy(g["decode_palette_norm"])
    got = ops.decode_palette(pred.to(dev), paln.to(dev)).cpu()
    assert got.dtype == torch.int64 and np.array_equal(got.numpy(), g["decode_out"])  # includes the tie case
    gen = torch.Generator().manual_seed(3)
    pred = torch.randn((2, 3, 896, 448), generator=gen)
    _, paln = glue_ref.create_palette(4, 2, train=True, generator=gen)
    want = glue_ref.process_pred_masks(pred, paln)
    assert torch.equal(ops.decode_palette(pred.to(dev), paln.to(dev)).cpu(), want)
    # fused cv2 INTER_NEAREST resize (src/predict.py:258) + nodata zeroing (src/predict_no_prompt.py:303)
    for crop in (512, 112, 336):
        nod = torch.from_numpy(np.random.default_rng(crop).random((2, crop, crop)) < 0.1)
        got = ops.decode_palette(pred.to(dev), paln.to(dev), out_size=crop, nodata=nod.to(dev), dtype=torch.uint8)
        for b in range(2):
            pm, _ = glue_ref.predict_tail(want[b].numpy(), crop)
            pm = pm.copy()
            pm[nod[b].numpy()] = 0
            assert np.array_equal(got[b].cpu().numpy(), pm.astype(np.uint8))


def test_vote_stitch_bit_exact(dev, g):
    counter = torch.zeros((30, 44), dtype=torch.int32, device=dev)
    boxes = torch.from_numpy(g["vote_boxes"]).to(dev)
    cls = torch.from_numpy(g["vote_cls"]).to(dev)
    ops.vote_accumulate(counter, cls, boxes, overlapping=True)  # all six tiles (two identical boxes) in ONE launch
    assert np.array_equal(counter.cpu().numpy().view(np.uint8).reshape(30, 44, 4), g["vote_counter"])
    assert np.array_equal(ops.vote_argmax(counter).cpu().numpy(), g["vote_argmax"])
    # tile-at-a-time (non-atomic path) gives the same canvas
    c2 = torch.zeros((30, 44), dtype=torch.int32, device=dev)
    for i in range(len(boxes)):
        ops.vote_accumulate(c2, cls[i:i + 1], boxes[i:i + 1])
    assert torch.equal(c2, counter)


def test_vote_uint8_wraparound(dev):
    """The reference counter is uint8 and wraps at 256 (src/predict.py:114-118): 300 votes -> 44, no carry."""
    counter = torch.zeros((8, 8), dtype=torch.int32, device=dev)
    n = 300
    cls = torch.full((n, 8, 8), 1, dtype=torch.uint8, device=dev)
    cls[:, :, 4:] = 3
    boxes = torch.tensor([[0, 0, 8, 8]] * n, dtype=torch.int32, device=dev)
    ops.vote_accumulate(counter, cls, boxes, overlapping=True)
    got = counter.cpu().numpy().view(np.uint8).reshape(8, 8, 4)
    want = np.zeros((8, 8, 4), dtype=np.uint8)
    want[:, :4, 1] = n % 256
    want[:, 4:, 3] = n % 256
    assert np.array_equal(got, want)


def test_full_scene_stitch_properties(dev):
    """BASELINE config 3 geometry (8000x4000 scene, 512 px tiles, stride 448 -> 18x9 = 162 tiles): the canvas equals
    the numpy Accumulator restatement, every pixel received between 1 and 4 votes, argmax is idempotent."""
    Hs, Ws, crop = 4000, 8000, 512
    boxes_np = synth.sliding_boxes(Hs, Ws, crop, 448)
    assert len(boxes_np) == 162
    rng = np.random.default_rng(0)
    cls_np = rng.integers(0, 4, size=(len(boxes_np), crop, crop)).astype(np.uint8)
    counter = torch.zeros((Hs, Ws), dtype=torch.int32, device=dev)
    ops.vote_accumulate(counter, torch.from_numpy(cls_np).to(dev), torch.from_numpy(boxes_np).to(dev))
    acc = glue_ref.AccumulatorRef((Hs, Ws))
    for b, c in zip(boxes_np, cls_np):
        acc.update(tuple(int(v) for v in b), np.eye(4, dtype=np.uint8)[c])
    got = counter.cpu().numpy().view(np.uint8).reshape(Hs, Ws, 4)
    assert np.array_equal(got, acc.counter)
    votes = got.sum(axis=2)
    assert votes.min() >= 1 and votes.max() <= 4
    assert np.array_equal(ops.vote_argmax(counter).cpu().numpy(), acc.argmax().astype(np.uint8))


@pytest.mark.parametrize("B", [1, 3])
def test_loss_golden_and_grad(dev, g, B):
    pred, lab, yes = (torch.from_numpy(g[f"loss_{k}_B{B}"]) for k in ("pred", "labels", "yes"))
    got = ops.smooth_l1_loss(pred.to(dev), lab.to(dev), yes.to(dev), 0.01, per_sample=False)
    np.testing.assert_allclose(got.item(), g[f"loss_out_B{B}"], rtol=2e-6)
    for per_sample in (False, True):
        p = pred.clone().requires_grad_(True)
        want = glue_ref.seggpt_loss(p, lab, yes, 0.01, per_sample=per_sample)
        want.backward()
        loss, grad = ops.smooth_l1_loss(pred.to(dev), lab.to(dev), yes.to(dev), 0.01, per_sample=per_sample,
                                        want_grad=True)
        np.testing.assert_allclose(loss.item(), want.item(), rtol=2e-6)
        np.testing.assert_allclose(grad.cpu().numpy(), p.grad.numpy(), rtol=1e-5, atol=1e-9)


def test_loss_full_size(dev):
    gen = torch.Generator().manual_seed(8)
    pred = torch.randn((2, 3, 896, 448), generator=gen)
    lab = torch.randn((2, 3, 448, 448), generator=gen)
    yes = torch.rand((2, 1, 448, 448), generator=gen) > 0.25
    for per_sample in (False, True):
        want = glue_ref.seggpt_loss(pred, lab, yes, 0.01, per_sample=per_sample)
        got = ops.smooth_l1_loss(pred.to(dev), lab.to(dev), yes.to(dev), 0.01, per_sample=per_sample)
        np.testing.assert_allclose(got.item(), want.item(), rtol=1e-5)


def _ingest_oracle(scene, nodata, boxes, crop):
    u8 = glue_ref.tif_image_4band(scene.astype(np.float32), nodata)
    imgs, crops, nds = [], [], []
    for b in boxes:
        ci, cn, _ = glue_ref.crop_tif(tuple(int(v) for v in b), u8, nodata, None, crop)
        crops.append(ci)
        nds.append(cn)
        imgs.append(glue_ref.normalize(torch.from_numpy(glue_ref.get_crop_image(ci, 448))[None])[0])
    return np.stack(crops), np.stack(nds), torch.stack(imgs)


@pytest.mark.parametrize("crop,Hs,Ws", [(512, 1100, 1300), (112, 300, 420), (448, 900, 1000), (1024, 1500, 2100)])
def test_ingest_bit_exact(dev, crop, Hs, Ws):
    scene = synth.scene_u16(Hs, Ws, seed=crop)
    nodata = synth.nodata_wedge(Hs, Ws)
    boxes = np.array([[0, 0, crop, crop], [Ws - crop - 3, Hs - crop - 5, Ws - 3, Hs - 5],
                      [-(crop // 3), -(crop // 4), crop - crop // 3, crop - crop // 4],
                      [Ws - crop // 2, Hs - crop // 2, Ws + crop - crop // 2, Hs + crop - crop // 2]], dtype=np.int32)
    want_u8, want_nd, want_img = _ingest_oracle(scene, nodata, boxes, crop)
    sc = torch.from_numpy(scene.view(np.int16)).to(dev)
    nd = torch.from_numpy(nodata).to(dev)
    stats = ops.scene_stats(sc, nd)
    comp = np.stack([scene[3], scene[2], scene[:2].astype(np.float32).mean(axis=0)]).astype(np.float32)
    want_stats = np.array([comp[:, ~nodata].min(), comp[0].max(), comp[1].max(), comp[2].max()], dtype=np.float32)
    assert np.array_equal(stats.cpu().numpy(), want_stats)
    out = ops.ingest_tiles(sc, nd, stats, torch.from_numpy(boxes).to(dev), crop, want_u8=True, want_nodata=True)
    assert np.array_equal(out["u8"].cpu().numpy(), want_u8)
    assert np.array_equal(out["nodata"].cpu().numpy().astype(bool), want_nd.astype(bool))
    assert torch.equal(out["image"].cpu(), want_img)


def _mosaic_inputs(N, H, W, seed):
    rng = np.random.default_rng(seed)
    data = (rng.random((N, 4, H, W), dtype=np.float32) * 4000.0 - 150.0).astype(np.float32)   # negative overshoot too
    yes = (rng.random((N, H, W)) > 0.35).astype(np.uint8) * 255                               # rasterio read_masks
    yes[:, : H // 7, : W // 5] = 0                                                           # a region no raster covers
    return data, yes


@pytest.mark.parametrize("N,H,W", [(1, 64, 80), (3, 301, 517), (5, 700, 900)])
def test_merge_mosaic_bit_exact(dev, N, H, W):
    """bseg_merge_mosaic vs the numpy restatement of merge_tifs' accumulation (src/util/geo_util.py:410-419)."""
    data, yes = _mosaic_inputs(N, H, W, seed=N)
    want_mean, want_mask = glue_ref.merge_mosaic(data, yes)
    mean, mask = ops.merge_mosaic(torch.from_numpy(data).to(dev), torch.from_numpy(yes).to(dev))
    assert np.array_equal(mask.cpu().numpy(), want_mask)
    assert np.array_equal(mean.cpu().numpy(), want_mean)
    # 0/1 weights give the same mean up to rounding of the *255 products; nodata is identical
    mean01, mask01 = ops.merge_mosaic(torch.from_numpy(data).to(dev), torch.from_numpy(yes // 255).to(dev))
    assert torch.equal(mask01, mask)
    np.testing.assert_allclose(mean01.cpu().numpy(), want_mean, rtol=2e-6, atol=1e-3)


@pytest.mark.parametrize("crop,Hs,Ws", [(512, 1100, 1300), (160, 300, 420)])
def test_ingest_f32_mosaic_bit_exact(dev, crop, Hs, Ws):
    """merge -> tif_image -> crop -> resize -> normalise on a float32 mosaic with negative values
    (src/util/geo_util.py:410-420,454-468; src/data.py:93-124)."""
    data, yes = _mosaic_inputs(3, Hs, Ws, seed=crop)
    scene, nodata = glue_ref.merge_mosaic(data, yes)
    assert scene[:, ~nodata].min() < 0
    boxes = np.array([[0, 0, crop, crop], [Ws - crop - 3, Hs - crop - 5, Ws - 3, Hs - 5],
                      [-(crop // 3), -(crop // 4), crop - crop // 3, crop - crop // 4]], dtype=np.int32)
    want_u8, want_nd, want_img = _ingest_oracle(scene, nodata, boxes, crop)
    sc, nd = ops.merge_mosaic(torch.from_numpy(data).to(dev), torch.from_numpy(yes).to(dev))
    stats = ops.scene_stats(sc, nd)
    comp = np.stack([scene[3], scene[2], scene[:2].mean(axis=0)])
    want_stats = np.array([comp[:, ~nodata].min(), comp[0].max(), comp[1].max(), comp[2].max()], dtype=np.float32)
    assert np.array_equal(stats.cpu().numpy(), want_stats)
    out = ops.ingest_tiles(sc, nd, stats, torch.from_numpy(boxes).to(dev), crop, want_u8=True, want_nodata=True)
    assert np.array_equal(out["u8"].cpu().numpy(), want_u8)
    assert np.array_equal(out["nodata"].cpu().numpy().astype(bool), want_nd.astype(bool))
    assert torch.equal(out["image"].cpu(), want_img)


@pytest.mark.parametrize("kind", ["u16", "f32"])
def test_scene_stats_of_row_ranges_merge_exactly(dev, kind):
    """bseg_scene_stats_rows / _finalize (the scene-global min / max of src/util/geo_util.py:459-464 when the rows of a
    scene live on several GPUs): keys of disjoint row ranges, merged with integer min / max as the all-reduce of
    ops.scene_stats_sharded does, decode to exactly the whole-scene statistics; rows outside a range are never read
    (they hold garbage here); an all-nodata range does not disturb the minimum."""
    Hs, Ws = 301, 517
    rng = np.random.default_rng(5)
    if kind == "u16":
        scene = synth.scene_u16(Hs, Ws, seed=3)
        sc_full = torch.from_numpy(scene.view(np.int16)).to(dev)
    else:
        scene = (rng.standard_normal((4, Hs, Ws)) * 900 + 300).astype(np.float32)
        sc_full = torch.from_numpy(scene).to(dev)
    nodata = rng.random((Hs, Ws)) < 0.2
    nodata[100:180] = True                                   # the middle range has no valid pixel at all
    nd = torch.from_numpy(nodata).to(dev)
    want = ops.scene_stats(sc_full, nd)
    assert torch.equal(ops.scene_stats_sharded(sc_full, nd, 0, Hs), want)   # world size 1: no merge
    L = _lib.lib()
    merged = None
    for r0, r1 in [(0, 100), (100, 180), (180, Hs), (Hs, Hs)]:
        part = sc_full.clone()
        garbage = torch.full_like(part, 30000 if kind == "u16" else 1e30)
        part[:, :r0], part[:, r1:] = garbage[:, :r0], garbage[:, r1:]
        keys = torch.empty(4, dtype=torch.int32, device=dev)
        _lib.check(L.bseg_scene_stats_rows(_lib.ptr(part), int(kind == "f32"), _lib.ptr(nd.to(torch.uint8)), Hs, Ws, r0,
                                           r1, _lib.ptr(keys), _lib.stream_ptr()))
        k = keys.to(torch.int64) & 0xFFFFFFFF
        merged = k if merged is None else torch.cat([torch.minimum(merged[:1], k[:1]), torch.maximum(merged[1:], k[1:])])
    stats = torch.empty(4, dtype=torch.float32, device=dev)
    _lib.check(L.bseg_scene_stats_finalize(_lib.ptr(merged.to(torch.int32)), _lib.ptr(stats), _lib.stream_ptr()))
    assert torch.equal(stats, want)
    with pytest.raises(_lib.BsegError):
        _lib.check(L.bseg_scene_stats_rows(_lib.ptr(sc_full), 0, _lib.ptr(nd.to(torch.uint8)), Hs, Ws, 5, Hs + 1,
                                           _lib.ptr(torch.empty(4, dtype=torch.int32, device=dev)), _lib.stream_ptr()))


@pytest.mark.parametrize("H,W", [(4, 256), (97, 131), (1000, 2001)])
def test_overlay_prediction_bit_exact(dev, H, W):
    """bseg_overlay_prediction vs the Pillow-pinned restatement (src/util/img_util.py:98-116); odd pixel counts
    exercise the scalar tail, class id 200 (outside the table) must copy the base pixel."""
    classes = ("nodata", "sand", "water", "veg")
    rng = np.random.default_rng(H)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    pred = rng.integers(0, 4, (H, W), dtype=np.uint8)
    if H == 4:  # every (grey level, class) pair
        img = np.repeat(np.arange(256, dtype=np.uint8)[None, :, None], 4, axis=0).repeat(3, axis=2).copy()
        pred = np.repeat(np.arange(4, dtype=np.uint8)[:, None], 256, axis=1).copy()
    want = glue_ref.overlay_prediction(img, pred, classes)
    got = ops.overlay_prediction(torch.from_numpy(img).to(dev), torch.from_numpy(pred).to(dev), classes)
    assert np.array_equal(got.cpu().numpy(), want)
    pred[0, :7] = 200
    want = glue_ref.overlay_prediction(img, pred, classes)
    got = ops.overlay_prediction(torch.from_numpy(img).to(dev), torch.from_numpy(pred).to(dev), classes)
    assert np.array_equal(got.cpu().numpy(), want)
    # a different class order, as conf.classes allows
    other = ("nodata", "water", "veg", "sand")
    got = ops.overlay_prediction(torch.from_numpy(img).to(dev), torch.from_numpy(pred).to(dev), other)
    assert np.array_equal(got.cpu().numpy(), glue_ref.overlay_prediction(img, pred, other))


def test_accumulator_image_paste_and_outputs(dev, tmp_path):
    """Accumulator.update with img_crop + save_current (src/predict.py:96-112,157): pasted canvas, overlay PNG and mask
    PNG equal the numpy/Pillow restatement."""
    from PIL import Image
    from beach_seg_b200.predict import Accumulator

    Hs, Ws, crop = 300, 420, 128
    classes = ("nodata", "sand", "water", "veg")
    rng = np.random.default_rng(3)
    scene_img = rng.integers(0, 256, (Hs, Ws, 3), dtype=np.uint8)
    boxes = synth.sliding_boxes(Hs, Ws, crop, 96)
    canvas = np.zeros((Hs, Ws, 3), dtype=np.uint8)
    ref = glue_ref.AccumulatorRef((Hs, Ws))
    with Accumulator((Hs, Ws), tmp_path, classes=classes, device=dev) as acc:
        for b in boxes:
            b = tuple(int(v) for v in b)
            ci = glue_ref.padded_crop(scene_img, *b, crop)
            cls = rng.integers(0, 4, (crop, crop)).astype(np.uint8)
            one_hot = np.eye(4, dtype=np.uint8)[cls]
            acc.update("20240101", b, one_hot, ci, None)
            ref.update(b, one_hot)
            x0, y0, x1, y1 = max(b[0], 0), max(b[1], 0), min(b[2], Ws), min(b[3], Hs)
            canvas[y0:y1, x0:x1] = ci[y0 - b[1]:y1 - b[1], x0 - b[0]:x1 - b[0]]
        assert np.array_equal(acc.current_img.cpu().numpy(), canvas)
        want_pred = ref.argmax().astype(np.uint8)
        assert np.array_equal(acc.overlay().cpu().numpy(), glue_ref.overlay_prediction(canvas, want_pred, classes))
    png = np.array(Image.open(tmp_path / "images" / "20240101.png"))
    assert np.array_equal(png, glue_ref.overlay_prediction(canvas, want_pred, classes))
    mask_png = np.array(Image.open(tmp_path / "masks" / "20240101.png"))
    assert np.array_equal(mask_png, want_pred)
