"""GPU: the SegGPT forward through the drop-in module (C ABI underneath) against the HF module (fp32, CPU) with the
same seeded weights.  north_star tolerance for bf16 operands: logits within 1e-2 relative; class maps identical
except pixels whose decision margin is within tolerance (count reported)."""
import numpy as np
import pytest
import torch

from beach_seg_b200 import ops, synth
from beach_seg_b200.seggpt import SegGptB200
from oracle import glue_ref
from oracle.seggpt_ref import make_reference_model, seggpt_forward

pytestmark = pytest.mark.gpu

SMALL = dict(num_layers=5, merge_index=1, intermediate=(1, 2, 3, 4))
REL_TOL = 1e-2  # north_star: "logits within 1e-2 relative in bf16"


def rel_l2(a, b):
    return ((a - b).norm() / b.norm()).item()


@pytest.fixture(scope="module")
def small(dev):
    hf = make_reference_model(seed=1, stress=True, **SMALL)
    return hf, SegGptB200.from_hf(hf, device=dev)


def test_small_model_layerwise(dev, small):
    """5-layer stress-initialised backbone: first localise any error (embeddings, per-layer residual stream, decoder
    input) with the restatement's intermediates, then check pred_masks."""
    hf, model = small
    px, ppx, pm = synth.model_inputs(batch=2, seed=9)
    cap = {}
    with torch.no_grad():
        want = seggpt_forward(hf.state_dict(), px, ppx, pm, capture=cap, **SMALL)
        got = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                    embedding_type="instance").pred_masks.cpu()
    r = rel_l2(got, want)
    bottom = rel_l2(got[:, :, 448:], want[:, :, 448:])
    print(f"[small model] pred rel-L2={r:.3e} bottom-half rel-L2={bottom:.3e} "
          f"max|err|/max|ref|={(got - want).abs().max().item() / want.abs().max().item():.3e}")
    assert got.shape == (2, 3, 896, 448)
    assert r < REL_TOL and bottom < REL_TOL


@pytest.mark.parametrize("embedding_type", ["instance", "semantic"])
def test_small_model_vs_hf(dev, small, embedding_type):
    hf, model = small
    px, ppx, pm = synth.model_inputs(batch=3, seed=21)
    with torch.no_grad():
        want = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type=embedding_type).pred_masks
        got = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                    embedding_type=embedding_type).pred_masks.cpu()
    assert rel_l2(got, want) < REL_TOL


def test_small_model_feature_ensemble(dev, small):
    hf, model = small
    px, ppx, pm = synth.model_inputs(batch=2, seed=10)
    px = px[:1].expand(2, -1, -1, -1).contiguous()
    with torch.no_grad():
        want = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance",
                  feature_ensemble=True).pred_masks
        got = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                    embedding_type="instance", feature_ensemble=True).pred_masks.cpu()
        # two tiles of two prompts each in one launch == two independent HF calls
        px4 = torch.cat([px, synth.model_inputs(2, seed=11)[0][:1].expand(2, -1, -1, -1)])
        ppx4, pm4 = torch.cat([ppx, ppx]), torch.cat([pm, pm])
        got4 = model(pixel_values=px4.to(dev), prompt_pixel_values=ppx4.to(dev), prompt_masks=pm4.to(dev),
                     feature_ensemble=True, ensemble_group=2).pred_masks.cpu()
        want_b = hf(pixel_values=px4[2:], prompt_pixel_values=ppx, prompt_masks=pm, feature_ensemble=True).pred_masks
    assert rel_l2(got, want) < REL_TOL
    assert rel_l2(got4[:2], want) < REL_TOL and rel_l2(got4[2:], want_b) < REL_TOL


def test_query_half_only_is_bit_identical(dev, small):
    """bseg_forward_query_half: the bottom half of pred_masks (the only part the reference reads, src/model.py:158-160)
    is bit-identical to the full forward, the top half is zero; also with the feature ensemble."""
    _, model = small
    px, ppx, pm = synth.model_inputs(batch=3, seed=33)
    kw = dict(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev))
    with torch.no_grad():
        full = model(**kw).pred_masks
        half = model(**kw, query_half_only=True).pred_masks
        assert torch.equal(half[:, :, 448:], full[:, :, 448:])
        assert not half[:, :, :448].any()
        full_e = model(**kw, feature_ensemble=True).pred_masks
        half_e = model(**kw, feature_ensemble=True, query_half_only=True).pred_masks
        assert torch.equal(half_e[:, :, 448:], full_e[:, :, 448:])


def test_interface_errors(dev, small):
    _, model = small
    z = torch.zeros((1, 3, 448, 448), device=dev)
    with pytest.raises(ValueError):
        model(pixel_values=torch.zeros((1, 3, 512, 512), device=dev), prompt_pixel_values=z, prompt_masks=z)
    with pytest.raises(ValueError):
        model(pixel_values=torch.zeros((1, 4, 448, 448), device=dev), prompt_pixel_values=z, prompt_masks=z)
    with pytest.raises(ValueError):
        model(pixel_values=z, prompt_pixel_values=z, prompt_masks=z, embedding_type="panoptic")
    out = model(pixel_values=z, prompt_pixel_values=z, prompt_masks=z)
    out.pred_masks = out.pred_masks.mean(dim=0, keepdim=True)  # assignable like src/predict_no_prompt.py:298
    assert out.loss is None and model.device == dev


def test_full_model_vs_hf_and_golden(dev, golden_dir):
    """24-layer ViT-L, stress init (all bias / affine / rel-pos paths active), batch 1 (BASELINE config 1 shape)."""
    g = np.load(golden_dir / "seggpt_golden.npz")
    hf = make_reference_model(seed=0, stress=True)
    model = SegGptB200.from_hf(hf, device=dev)
    px, ppx, pm = synth.model_inputs(batch=1, seed=123)
    with torch.no_grad():
        want = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance").pred_masks
        got = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                    embedding_type="instance").pred_masks.cpu()
    # the box regenerates the same reference as the committed fixture
    np.testing.assert_allclose(want[:, :, ::16, ::16].numpy(), g["stress_slice"], rtol=0, atol=5e-5)
    r, rb = rel_l2(got, want), rel_l2(got[:, :, 448:], want[:, :, 448:])
    mx = (got - want).abs().max().item() / want.abs().max().item()
    # class-map agreement (fixed palette build_palette(3), as in the reference's non-random branch)
    _, paln = glue_ref.create_palette(4, 1, train=False)
    cls_ref = glue_ref.process_pred_masks(want, paln)
    cls_got = ops.decode_palette(got.to(dev), paln.to(dev)).cpu()
    flips = (cls_ref != cls_got)
    # decision margin (d2 - d1) of the reference at the flipped pixels
    H = 448
    d = ((want[0, :, H:].permute(1, 2, 0)[:, :, None, :] - paln[0][None, None]) ** 2).sum(-1)
    top2 = d.topk(2, dim=-1, largest=False).values
    margin = (top2[..., 1] - top2[..., 0])
    fm = margin[flips[0]]
    print(f"[full model] rel-L2={r:.3e} bottom rel-L2={rb:.3e} max|err|/max|ref|={mx:.3e} "
          f"class flips={int(flips.sum())}/{flips.numel()} max flipped margin={fm.max().item() if fm.numel() else 0:.3e}")
    assert r < REL_TOL and rb < REL_TOL
    assert flips.float().mean().item() < 0.01
    if fm.numel():
        assert fm.max().item() < 0.1  # only pixels whose logit sits on the decision boundary may flip


def _decision_margin(pred_ref: torch.Tensor, paln: torch.Tensor) -> torch.Tensor:
    """(d2 - d1) of the reference's palette decode (src/model.py:155-175) per query pixel of sample 0: [448, 448]."""
    d = ((pred_ref[0, :, 448:].permute(1, 2, 0)[:, :, None, :] - paln[0][None, None]) ** 2).sum(-1)
    top2 = d.topk(2, dim=-1, largest=False).values
    return top2[..., 1] - top2[..., 0]


@pytest.mark.parametrize("B", [1, 3])
def test_hf_loss_value_matches_hf(dev, small, B):
    """`out.loss` when `labels=` is passed (src/model.py:245-251 does; HF:modeling_seggpt.py:780-819,910-917).  HF's
    default bool_masked_pos has batch dimension 1, so its loss is the sum over all B samples divided by ONE sample's
    mask count (B x the per-sample mean); an explicit [B, 1568] mask divides by B x as much."""
    hf, model = small
    px, ppx, pm = synth.model_inputs(batch=B, seed=70 + B)
    labels = synth.model_inputs(batch=B, seed=80 + B)[2]
    bmp = torch.cat([torch.zeros(784, dtype=torch.bool), torch.ones(784, dtype=torch.bool)])[None].repeat(B, 1)
    with torch.no_grad():
        want = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, labels=labels)
        want_b = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, labels=labels, bool_masked_pos=bmp)
        kw = dict(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                  labels=labels.to(dev))
        got = model(**kw)
        got_b = model(**kw, bool_masked_pos=bmp)
        # the loss HF would compute on OUR pred_masks (isolates the loss arithmetic from the bf16 forward deviation)
        from transformers.models.seggpt.modeling_seggpt import SegGptLoss as HfLoss

        on_ours = HfLoss(hf.config)(pm, got.pred_masks.cpu(), labels, bmp[:1])
    print(f"[HF loss B={B}] ours {got.loss.item():.6f} HF {want.loss.item():.6f} HF-on-our-pred {on_ours.item():.6f} | "
          f"explicit mask: ours {got_b.loss.item():.6f} HF {want_b.loss.item():.6f}")
    assert abs(got.loss.item() - on_ours.item()) <= 1e-5 * abs(on_ours.item())
    assert abs(got.loss.item() - want.loss.item()) <= 1e-2 * abs(want.loss.item())
    assert abs(got_b.loss.item() - want_b.loss.item()) <= 1e-2 * abs(want_b.loss.item())
    if B > 1:
        assert abs(got.loss.item() / got_b.loss.item() - B) < 1e-4
    # training path (graph to the prompt): same value
    p = ppx.to(dev).requires_grad_(True)
    out_t = model(pixel_values=px.to(dev), prompt_pixel_values=p, prompt_masks=pm.to(dev), labels=labels.to(dev))
    assert abs(out_t.loss.item() - got.loss.item()) <= 1e-6 * abs(got.loss.item())


def test_loss_scalar_is_bit_reproducible(dev):
    """The smooth-L1 reduction is two-stage in a fixed order (no float atomics): repeated launches give the same bits,
    for the loss as for the gradient."""
    g = torch.Generator().manual_seed(5)
    pred = torch.randn((8, 3, 896, 448), generator=g).to(dev)
    lab = torch.randn((8, 3, 448, 448), generator=g).to(dev)
    yes = (torch.rand((8, 448, 448), generator=g) > 0.25).to(dev)
    first = None
    for _ in range(5):
        loss, grad = ops.smooth_l1_loss(pred, lab, yes, 0.01, per_sample=False, want_grad=True)
        bits = (loss.view(torch.int32).item(), grad.view(torch.int32).sum().item())
        first = first or bits
        assert bits == first


def test_batch64_launch_tiles_0_and_63_vs_hf(dev):
    """The bench configuration itself: ONE 24-layer launch of 64 samples (BASELINE configs[1]); samples 0 and 63 are
    compared with the HF fp32 CPU oracle (2 CPU forwards)."""
    hf = make_reference_model(seed=0, stress=True)
    model = SegGptB200.from_hf(hf, device=dev, max_batch=64)
    px, ppx, pm = synth.model_inputs(batch=64, seed=640)
    with torch.no_grad():
        got = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                    embedding_type="instance").pred_masks
        _, paln = glue_ref.create_palette(4, 1, train=False)
        for i in (0, 63):
            want = hf(pixel_values=px[i:i + 1], prompt_pixel_values=ppx[i:i + 1], prompt_masks=pm[i:i + 1],
                      embedding_type="instance").pred_masks
            g = got[i:i + 1].cpu()
            r = rel_l2(g, want)
            flipped = ops.decode_palette(got[i:i + 1], paln.to(dev)).cpu() != glue_ref.process_pred_masks(want, paln)
            fm = _decision_margin(want, paln)[flipped[0]]
            print(f"[batch-64 launch] sample {i}: rel-L2={r:.3e} flips={int(flipped.sum())} "
                  f"max flipped margin={fm.max().item() if fm.numel() else 0:.3e}")
            assert r < REL_TOL
            assert flipped.float().mean().item() < 0.01
            if fm.numel():
                assert fm.max().item() < 0.1


def test_reference_written_prompt_batch_loads_and_predicts(dev, golden_dir, tmp_path, small):
    """`prompt_batch.pt` written by the reference's OWN create_trainable_params + handle_item +
    torch.save (tests/golden/ref_prompt_batch.pt.gz, oracle/make_golden_train.py) is loaded the way src/predict.py:213-216
    does, into a GPU PromptModel, which then predicts; checked against the oracle pipeline driven by the same file."""
    import gzip

    from beach_seg_b200 import train as artifacts
    from beach_seg_b200.config import BeachSegConfig
    from beach_seg_b200.model import PromptModel

    hf, backbone = small
    path = tmp_path / "prompt_batch.pt"
    path.write_bytes(gzip.open(golden_dir / "ref_prompt_batch.pt.gz").read())
    pm_model = PromptModel(BeachSegConfig(checkpoint="random-init:0"), device=dev, model=backbone)
    artifacts.load_prompt_batch(pm_model, path)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert len(pm_model.prompt_params_list) == len(raw["image"]) == 2
    assert all(torch.equal(a.detach().cpu(), b.detach()) for a, b in zip(pm_model.prompt_params_list, raw["image"]))
    px = synth.normalize(synth.smooth_image(2, seed=53))
    idx = torch.tensor([1, 0])
    torch.manual_seed(7)
    got = pm_model({"image": px.to(dev), "crop_idx": idx}).cpu()
    torch.manual_seed(7)
    pal, paln = glue_ref.create_palette(4, 2, train=True)
    ppx = glue_ref.normalize(torch.stack([raw["image"][i].detach() for i in idx.tolist()]))
    pmask = glue_ref.normalize(glue_ref.torch_apply_mask_rgb(pal, raw["mask"][idx]))
    with torch.no_grad():
        pred = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pmask, embedding_type="instance").pred_masks
    want = glue_ref.process_pred_masks(pred, paln)
    flips = (got != want).float().mean().item()
    print(f"[reference prompt_batch.pt] class-map flips vs oracle pipeline: {flips * 100:.3f} %")
    assert flips < 0.01
    # and the file this engine writes is read back by the reference's gather (src/model.py:189-192)
    artifacts.save_prompt_batch(pm_model, tmp_path / "ours")
    ours = torch.load(tmp_path / "ours" / "prompt_batch.pt", map_location="cpu", weights_only=False)
    assert set(ours) == set(raw) and all(torch.equal(a.detach(), b.detach()) for a, b in zip(ours["image"], raw["image"]))
    assert torch.equal(ours["mask"], raw["mask"]) and ours["date"] == raw["date"]


def test_prompt_model_forward_matches_reference_pipeline(dev):
    """PromptModel.forward (src/model.py:132-147) end to end: random palette from the global RNG, prompt gather,
    colourise, SegGPT, palette decode -- against the oracle restatement driven by the HF module (default init)."""
    from beach_seg_b200.config import BeachSegConfig
    from beach_seg_b200.model import PromptModel

    conf = BeachSegConfig(checkpoint="random-init:0")
    pm_model = PromptModel(conf, device=dev)
    prompt_img01 = synth.smooth_image(2, seed=50)            # prompt images in [0,1] (what get_crop returns)
    prompt_cls = synth.blocky_mask(2, seed=51)

    class DM:
        prompt_imgs = [{"image": prompt_img01[i], "mask": prompt_cls[i][None], "crop_idx": i} for i in range(2)]

    pm_model.create_trainable_params(DM)
    px = synth.normalize(synth.smooth_image(1, seed=52))
    batch = {"image": px.to(dev), "crop_idx": torch.tensor([1])}
    torch.manual_seed(99)
    got = pm_model(batch).cpu()

    hf = make_reference_model(seed=0, stress=False)
    torch.manual_seed(99)
    pal, paln = glue_ref.create_palette(4, 1, train=True)
    ppx = glue_ref.normalize(prompt_img01[1:2])
    pmask = glue_ref.normalize(glue_ref.torch_apply_mask_rgb(pal, prompt_cls[1:2][:, None]))
    with torch.no_grad():
        pred = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pmask, embedding_type="instance").pred_masks
    want = glue_ref.process_pred_masks(pred, paln)
    flipped = got != want
    flips = flipped.float().mean().item()
    fm = _decision_margin(pred, paln)[flipped[0]]
    print(f"[PromptModel.forward] class-map flips vs reference pipeline: {flips * 100:.3f} %, max flipped margin "
          f"{fm.max().item() if fm.numel() else 0:.3e}")
    assert got.shape == (1, 448, 448) and got.dtype == torch.int64
    assert flips < 0.01  # measured 0.2-0.3 %
    if fm.numel():
        assert fm.max().item() < 0.1  # only pixels on the reference's own decision boundary may flip


def test_full_scene_sliding_window_sharded_equals_single(dev):
    """BASELINE config 3 at full size: 8000x4000 uint16 scene, 512-px tiles, stride 448 (64-px overlap) = 18 x 9 = 162
    tiles, random-init 24-layer backbone.  Size-independent properties instead of a CPU oracle run (162 tiles would
    take ~15 min on the host): (1) every pixel's vote total equals the number of tiles covering it, (2) the canvas
    stitched from 8 owner-computes shards (no data-path collective, just the sum of the per-rank u32 canvases) is
    bit-identical to the single-rank canvas, (3) so is the final class map."""
    from beach_seg_b200.ml_util import load_model
    from beach_seg_b200.predict import TilePredictor, create_palette, shard_tiles

    Hs, Ws, crop = 4000, 8000, 512
    model = load_model("random-init:0", device=dev)
    predictor = TilePredictor(model, crop)
    scene = torch.from_numpy(synth.scene_u16(Hs, Ws, seed=7).view(np.int16)).to(dev)
    nodata = torch.from_numpy(synth.nodata_wedge(Hs, Ws, 0.05)).to(dev)
    boxes_np = synth.sliding_boxes(Hs, Ws, crop, 448)
    assert len(boxes_np) == 162
    boxes = torch.from_numpy(boxes_np).to(dev)
    stats = ops.scene_stats(scene, nodata)
    n = len(boxes_np)
    prompts = synth.normalize(synth.smooth_image(1, 2000)).to(dev).expand(n, -1, -1, -1).contiguous()
    pcls = synth.blocky_mask(1, 3000).to(dev).expand(n, -1, -1).contiguous()
    torch.manual_seed(42)
    palette = create_palette(4, n, True, dev)

    def run(ids):
        canvas = torch.zeros((Hs, Ws), dtype=torch.int32, device=dev)
        for s in range(0, len(ids), 64):
            sel = torch.as_tensor(list(ids[s:s + 64]), device=dev)
            cls = predictor.predict_tiles(scene, nodata, stats, boxes[sel], prompts[sel], pcls[sel],
                                          (palette[0][sel], palette[1][sel]))
            ops.vote_accumulate(canvas, cls, boxes[sel], overlapping=True)
        return canvas

    single = run(range(n))
    sharded = torch.zeros_like(single)
    for r in range(8):
        sharded += run(shard_tiles(n, r, 8))
    assert torch.equal(single, sharded)
    votes = single.cpu().numpy().view(np.uint8).reshape(Hs, Ws, 4).sum(axis=2)
    cover = np.zeros((Hs, Ws), dtype=np.int32)
    for x0, y0, x1, y1 in boxes_np:
        cover[max(y0, 0):min(y1, Hs), max(x0, 0):min(x1, Ws)] += 1
    assert np.array_equal(votes, cover)
    assert torch.equal(ops.vote_argmax(single), ops.vote_argmax(sharded))


def test_forward_and_gradient_bit_identical_with_cta_pair_gemm(dev, small):
    """Every GEMM epilogue of the forward (embedding table, QKV head split, residual, GELU, pixel shuffle) and of the
    backward (dgrad, GELU') through the CTA-pair kernel (tcgen05.mma.cta_group::2, 256-row tiles): same accumulation
    order as the one-CTA kernel, so pred_masks and the prompt gradient must be bit-identical."""
    from beach_seg_b200 import _lib

    hf, model = small
    px, ppx, pm = synth.model_inputs(batch=3, seed=33)
    L = _lib.lib()
    prev = L.bseg_gemm_set_cta_pairs(0)
    outs = []
    try:
        for pairs in (0, 1):
            L.bseg_gemm_set_cta_pairs(pairs)
            p = ppx.to(dev).requires_grad_(True)
            out = model(pixel_values=px.to(dev), prompt_pixel_values=p, prompt_masks=pm.to(dev),
                        embedding_type="instance").pred_masks
            d = torch.zeros_like(out)
            d[:, :, 448:] = torch.randn((3, 3, 448, 448), generator=torch.Generator().manual_seed(1)).to(dev)
            out.backward(d)
            torch.cuda.synchronize()
            outs.append((out.detach().clone(), p.grad.clone()))
    finally:
        L.bseg_gemm_set_cta_pairs(prev)
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
