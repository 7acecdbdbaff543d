"""GPU: native-resolution mode: 512-px tiles without the resize to 448, i.e. HF
`SegGptConfig(image_size=(1024, 512))` -- 64 x 32 tokens, T = 2048, rel-pos tables of 127 / 63 rows, position embeddings
interpolated 14x14 -> 64x32.  Oracle = that HF module (seeded random init, fp32, CPU); bar = north_star's 1e-2 relative
for bf16 operands, class maps differing only where the reference's decision margin is < 0.1."""
import pytest
import torch

from beach_seg_b200 import _lib, ops, synth
from beach_seg_b200.seggpt import SegGptB200
from oracle import glue_ref
from oracle.seggpt_ref import attention_ref, make_reference_model

pytestmark = pytest.mark.gpu
GH, GW = 64, 32
T = GH * GW
Q_SCALE = 0.125 * 1.4426950408889634
SMALL = dict(num_layers=5, merge_index=1, intermediate=(1, 2, 3, 4))


def rel_l2(a, b):
    return ((a - b).norm() / b.norm()).item()


def bf16r(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("nseq,qscale,relscale", [(1, 1.0, 0.3), (2, 3.0, 0.5)])
def test_attention_64x32_matches_oracle(dev, nseq, qscale, relscale):
    """bseg_attention_grid at the 64 x 32 token grid (key blocks of 64 = 2 token rows) against the oracle's attention."""
    g = torch.Generator().manual_seed(int(nseq * 100 + qscale * 10 + relscale * 7))
    qs = (torch.randn((nseq, 16, T, 64), generator=g) * qscale * Q_SCALE).to(torch.bfloat16)
    q = qs.float() / Q_SCALE
    k = bf16r(torch.randn((nseq, 16, T, 64), generator=g))
    v = bf16r(torch.randn((nseq, 16, T, 64), generator=g))
    rel_h = bf16r(torch.randn((2 * GH - 1, 64), generator=g) * relscale)
    rel_w = bf16r(torch.randn((2 * GW - 1, 64), generator=g) * relscale)
    want = attention_ref(q.reshape(-1, T, 64), k.reshape(-1, T, 64), v.reshape(-1, T, 64), rel_h, rel_w, GH, GW)
    want = want.reshape(nseq, 16, T, 64).permute(0, 2, 1, 3).reshape(nseq, T, 1024)
    L = _lib.lib()
    rows = L.bseg_relcat_rows(GH, GW)
    assert rows == 192
    relcat = torch.empty((rows, 64), dtype=torch.bfloat16, device=dev)
    rh, rw = rel_h.to(dev), rel_w.to(dev)
    _lib.check(L.bseg_pack_relcat_grid(_lib.ptr(rh), _lib.ptr(rw), _lib.ptr(relcat), GH, GW, _lib.stream_ptr()))
    qb, kb = qs.to(dev).contiguous(), k.to(dev).to(torch.bfloat16).contiguous()
    vt = v.to(dev).to(torch.bfloat16).transpose(2, 3).contiguous()
    out = torch.empty((nseq, T, 1024), dtype=torch.bfloat16, device=dev)
    lse = torch.empty((nseq, 16, T), dtype=torch.float32, device=dev)
    _lib.check(L.bseg_attention_grid(_lib.ptr(qb), _lib.ptr(kb), _lib.ptr(vt), _lib.ptr(relcat), _lib.ptr(out),
                                     _lib.ptr(lse), nseq, GH, GW, _lib.stream_ptr()), "bseg_attention_grid")
    torch.cuda.synchronize()
    got = out.float().cpu()
    rel = rel_l2(got, want)
    print(f"[attention 64x32 nseq={nseq} q*{qscale} rel*{relscale}] rel-L2={rel:.3e}")
    assert rel < 1e-2
    # the saved log-sum-exp (log2 domain) against the oracle's scores
    s = (q.reshape(-1, T, 64) * 0.125) @ k.reshape(-1, T, 64).transpose(-2, -1)
    from oracle.seggpt_ref import rel_pos_bias

    s = s + rel_pos_bias(q.reshape(-1, T, 64), rel_h, rel_w, GH, GW)
    lse_ref = torch.logsumexp(s, dim=-1) * 1.4426950408889634
    lse_err = (lse.cpu().reshape(-1, T) - lse_ref).abs().max().item()
    print(f"[attention 64x32] max |lse error| = {lse_err:.3e} at max |lse| = {lse_ref.abs().max().item():.1f} (log2 units)")
    assert lse_err < 2e-2 + 1e-3 * lse_ref.abs().max().item()  # bf16 operands: the scores themselves carry ~1e-3 relative


def test_unknown_token_grid_is_rejected(dev):
    L = _lib.lib()
    z = torch.zeros(16, dtype=torch.bfloat16, device=dev)
    rc = L.bseg_attention_grid(_lib.ptr(z), _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), None, 1, 96, 48,
                               _lib.stream_ptr())
    assert rc != 0 and b"not built" in L.bseg_last_error()


@pytest.mark.parametrize("embedding_type", ["instance", "semantic"])
def test_native_512_model_vs_hf(dev, embedding_type):
    """5-layer stress-initialised backbone at image_size 512: pred_masks [B, 3, 1024, 512] against the HF module."""
    hf = make_reference_model(seed=4, stress=True, image_size=512, **SMALL)
    model = SegGptB200.from_hf(hf, device=dev)
    assert model.image_size == 512 and model.num_patches == 2048
    px, ppx, pm = synth.model_inputs(batch=2, seed=31, size=512)
    with torch.no_grad():
        want = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type=embedding_type).pred_masks
        got = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                    embedding_type=embedding_type).pred_masks
        half = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                     embedding_type=embedding_type, query_half_only=True).pred_masks
    assert got.shape == (2, 3, 1024, 512)
    r = rel_l2(got.cpu(), want)
    print(f"[native 512 {embedding_type}] pred rel-L2={r:.3e}")
    assert r < 1e-2
    assert torch.equal(half[:, :, 512:], got[:, :, 512:]) and not half[:, :, :512].any()
    # class maps through the palette decode (generic in the tile size)
    _, paln = glue_ref.create_palette(4, 2, train=False)
    cls = ops.decode_palette(got, paln.to(dev)).cpu()
    d = ((want[:, :, 512:].permute(0, 2, 3, 1)[:, :, :, None, :] - paln[:, None, None]) ** 2).sum(-1)
    cls_ref = d.argmin(-1)
    top2 = d.topk(2, dim=-1, largest=False).values
    margin = top2[..., 1] - top2[..., 0]
    flipped = cls != cls_ref
    print(f"[native 512 {embedding_type}] class flips {int(flipped.sum())}/{flipped.numel()}")
    assert cls.shape == (2, 512, 512) and flipped.float().mean().item() < 0.01
    if flipped.any():
        assert margin[flipped].max().item() < 0.1


def test_native_512_interface(dev):
    """Wrong tile size raises HF's ValueError; the train step and the fp32 mode are 448-only and say so."""
    hf = make_reference_model(seed=4, stress=False, image_size=512, **SMALL)
    model = SegGptB200.from_hf(hf, device=dev)
    z448 = torch.zeros((1, 3, 448, 448), device=dev)
    with pytest.raises(ValueError):
        model(pixel_values=z448, prompt_pixel_values=z448, prompt_masks=z448)
    z = torch.zeros((1, 3, 512, 512), device=dev)
    with pytest.raises(NotImplementedError):
        model(pixel_values=z, prompt_pixel_values=z.clone().requires_grad_(True), prompt_masks=z)
    with pytest.raises(NotImplementedError):
        SegGptB200.from_hf(hf, device=dev, precision="fp32")


@pytest.mark.parametrize("f32", [False, True])
def test_native_ingest_bit_exact(dev, f32):
    """bseg_ingest_native_*: tif_image + crop_tif + /255 + Normalize with NO resize (src/data.py:94 skips it when
    inpt_size == crop_size), bit-exact against the oracle glue -- aligned boxes (vector path), unaligned and
    out-of-scene boxes (scalar path with the reference's zero / nodata padding)."""
    import numpy as np

    Hs, Ws, crop = 700, 1100, 512
    scene = synth.scene_u16(Hs, Ws, seed=21)
    nodata = synth.nodata_wedge(Hs, Ws, 0.05)
    boxes = np.array([[0, 0, 512, 512], [512, 128, 1024, 640], [301, 77, 813, 589], [-40, -60, 472, 452],
                      [900, 500, 1412, 1012]], dtype=np.int32)
    data = scene.astype(np.float32)
    u8 = glue_ref.tif_image_4band(data.copy(), nodata)
    sc = torch.from_numpy(data if f32 else scene.view(np.int16)).to(dev)
    nd = torch.from_numpy(nodata).to(dev)
    stats = ops.scene_stats(sc, nd)
    out = ops.ingest_tiles(sc, nd, stats, torch.from_numpy(boxes).to(dev), crop, want_u8=True, want_nodata=True,
                           out_size=crop)
    assert out["image"].shape == (len(boxes), 3, crop, crop)
    for i, b in enumerate(boxes):
        ci, cn, _ = glue_ref.crop_tif(tuple(int(v) for v in b), u8, nodata, None, crop)
        want = glue_ref.normalize(torch.from_numpy(glue_ref.get_crop_image(ci, crop))[None])[0]
        assert np.array_equal(out["u8"][i].cpu().numpy(), ci), i
        assert np.array_equal(out["nodata"][i].cpu().numpy().astype(bool), cn), i
        assert torch.equal(out["image"][i].cpu(), want), i


def test_native_tile_predictor_end_to_end(dev):
    """TilePredictor on a native-resolution backbone: ingest (no resize) -> forward at 512 -> palette decode at 512 ->
    votes, against the oracle pipeline driven by the HF module."""
    import numpy as np

    from beach_seg_b200.predict import TilePredictor

    hf = make_reference_model(seed=4, stress=True, image_size=512, **SMALL)
    model = SegGptB200.from_hf(hf, device=dev)
    with pytest.raises(ValueError):
        TilePredictor(model, 448)
    pred = TilePredictor(model, 512, random_palette=False)
    Hs, Ws = 512, 1024
    scene = synth.scene_u16(Hs, Ws, seed=22)
    nodata = np.zeros((Hs, Ws), dtype=bool)
    boxes = np.array([[0, 0, 512, 512], [512, 0, 1024, 512]], dtype=np.int32)
    sc, nd = torch.from_numpy(scene.view(np.int16)).to(dev), torch.from_numpy(nodata).to(dev)
    stats = ops.scene_stats(sc, nd)
    ppx = synth.normalize(synth.smooth_image(2, 60, size=512))
    pcls = synth.blocky_mask(2, 61, size=512)
    cls = pred.predict_tiles(sc, nd, stats, torch.from_numpy(boxes).to(dev), ppx.to(dev), pcls.to(dev)).cpu()
    assert cls.shape == (2, 512, 512) and cls.dtype == torch.uint8
    u8 = glue_ref.tif_image_4band(scene.astype(np.float32), nodata)
    pal, paln = glue_ref.create_palette(4, 2, train=False)
    px = torch.stack([glue_ref.normalize(torch.from_numpy(glue_ref.get_crop_image(
        glue_ref.crop_tif(tuple(int(v) for v in b), u8, nodata, None, 512)[0], 512))[None])[0] for b in boxes])
    pm = glue_ref.normalize(glue_ref.torch_apply_mask_rgb(pal, pcls[:, None]))
    with torch.no_grad():
        want = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance").pred_masks
    d = ((want[:, :, 512:].permute(0, 2, 3, 1)[:, :, :, None, :] - paln[:, None, None]) ** 2).sum(-1)
    flips = (cls.long() != d.argmin(-1)).float().mean().item()
    print(f"[native TilePredictor] class flips vs oracle pipeline {flips * 100:.3f} %")
    assert flips < 0.01


def test_attention_128x64_matches_oracle(dev):
    """bseg_attention_grid at the 128 x 64 token grid of native 1024-px tiles (T = 8192, key blocks of one token row, the
    rel tables of 255 + 127 rows go through TMEM in two passes) against the oracle's attention, evaluated on the GPU one
    head at a time (the 8192 x 8192 score matrix of a head is 268 MB)."""
    gh, gw = 128, 64
    t = gh * gw
    g = torch.Generator().manual_seed(17)
    qs = (torch.randn((1, 16, t, 64), generator=g) * 1.5 * Q_SCALE).to(torch.bfloat16).to(dev)
    q = qs.float() / Q_SCALE
    k = bf16r(torch.randn((1, 16, t, 64), generator=g)).to(dev)
    v = bf16r(torch.randn((1, 16, t, 64), generator=g)).to(dev)
    rel_h = bf16r(torch.randn((2 * gh - 1, 64), generator=g) * 0.4).to(dev)
    rel_w = bf16r(torch.randn((2 * gw - 1, 64), generator=g) * 0.4).to(dev)
    L = _lib.lib()
    rows = L.bseg_relcat_rows(gh, gw)
    assert rows == 384
    relcat = torch.empty((rows, 64), dtype=torch.bfloat16, device=dev)
    _lib.check(L.bseg_pack_relcat_grid(_lib.ptr(rel_h), _lib.ptr(rel_w), _lib.ptr(relcat), gh, gw, _lib.stream_ptr()))
    kb = k.to(torch.bfloat16).contiguous()
    vt = v.to(torch.bfloat16).transpose(2, 3).contiguous()
    out = torch.empty((1, t, 1024), dtype=torch.bfloat16, device=dev)
    _lib.check(L.bseg_attention_grid(_lib.ptr(qs), _lib.ptr(kb), _lib.ptr(vt), _lib.ptr(relcat), _lib.ptr(out), None, 1,
                                     gh, gw, _lib.stream_ptr()), "bseg_attention_grid")
    torch.cuda.synchronize()
    got = out.float().reshape(t, 16, 64)
    ih = (torch.arange(gh)[:, None] - torch.arange(gh)[None, :] + gh - 1).to(dev)
    iw = (torch.arange(gw)[:, None] - torch.arange(gw)[None, :] + gw - 1).to(dev)
    num = den = 0.0
    for hd in range(16):
        rq = q[0, hd].reshape(gh, gw, 64)
        bias = (torch.einsum("hwc,hkc->hwk", rq, rel_h[ih])[:, :, :, None] +
                torch.einsum("hwc,wkc->hwk", rq, rel_w[iw])[:, :, None, :]).reshape(t, t)
        want = torch.softmax((q[0, hd] * 0.125) @ k[0, hd].T + bias, dim=-1) @ v[0, hd]
        num += (got[:, hd] - want).pow(2).sum().item()
        den += want.pow(2).sum().item()
        del bias, want
    rel = (num / den) ** 0.5
    print(f"[attention 128x64] rel-L2={rel:.3e}")
    assert rel < 1e-2


def test_native_1024_model_vs_hf(dev):
    """4-layer stress-initialised backbone at image_size 1024 (T = 8192): pred_masks [1, 3, 2048, 1024] against the HF
    module (fp32, CPU: about 20 GB of score matrices, one sample)."""
    # (HF concatenates one hidden state per DISTINCT intermediate index and the decoder needs four: four layers)
    hf = make_reference_model(seed=6, stress=True, image_size=1024, num_layers=4, merge_index=0, intermediate=(0, 1, 2, 3))
    model = SegGptB200.from_hf(hf, device=dev)
    assert model.image_size == 1024 and model.num_patches == 8192
    px, ppx, pm = synth.model_inputs(batch=1, seed=41, size=1024)
    with torch.no_grad():
        want = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance").pred_masks
        got = model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev),
                    embedding_type="instance").pred_masks
    assert got.shape == (1, 3, 2048, 1024)
    r = rel_l2(got.cpu(), want)
    print(f"[native 1024] pred rel-L2={r:.3e}")
    assert r < 1e-2
