"""GPU: edge cases of the hot path -- empty and ragged batches, tiles that leave the scene, all-nodata scenes, argument
errors of the C ABI (same failure behaviour through the Python mirror: an exception, never a silent fallback)."""
import ctypes as C

import numpy as np
import pytest
import torch

from beach_seg_b200 import _lib, ops, synth
from beach_seg_b200.predict import Accumulator, TilePredictor, create_palette, shard_tiles
from beach_seg_b200.seggpt import SegGptB200
from oracle import glue_ref
from oracle.seggpt_ref import make_reference_model

pytestmark = pytest.mark.gpu
SMALL = dict(num_layers=5, merge_index=1, intermediate=(1, 2, 3, 4))


@pytest.fixture(scope="module")
def small_model(dev):
    return SegGptB200.from_hf(make_reference_model(seed=1, stress=True, **SMALL), device=dev, max_batch=2)


def test_ragged_batch_equals_one_launch(dev, small_model):
    """max_batch=2: a batch of 5 runs as 2+2+1 and must equal the per-sample results bit for bit."""
    px, ppx, pm = synth.model_inputs(batch=5, seed=4)
    with torch.no_grad():
        got = small_model(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev)).pred_masks
        for i in (0, 4):
            one = small_model(pixel_values=px[i:i + 1].to(dev), prompt_pixel_values=ppx[i:i + 1].to(dev),
                              prompt_masks=pm[i:i + 1].to(dev)).pred_masks
            assert torch.equal(one[0], got[i])


def test_empty_tile_batch_and_empty_shard(dev, small_model):
    crop = 128
    scene = torch.from_numpy(synth.scene_u16(256, 256, seed=2).view(np.int16)).to(dev)
    nodata = torch.zeros((256, 256), dtype=torch.bool, device=dev)
    stats = ops.scene_stats(scene, nodata)
    pred = TilePredictor(small_model, crop)
    none = torch.zeros((0, 4), dtype=torch.int32, device=dev)
    out = pred.predict_tiles(scene, nodata, stats, none, torch.zeros((0, 3, 448, 448), device=dev),
                             torch.zeros((0, 448, 448), dtype=torch.uint8, device=dev))
    assert out.shape == (0, crop, crop)
    tiles = ops.ingest_tiles(scene, nodata, stats, none, crop, want_u8=True, want_nodata=True)
    assert tiles["image"].shape == (0, 3, 448, 448) and tiles["u8"].shape == (0, crop, crop, 3)
    canvas = torch.zeros((256, 256), dtype=torch.int32, device=dev)
    ops.vote_accumulate(canvas, out, none)
    assert not canvas.any()
    assert len(shard_tiles(3, 7, 8)) == 0 and sum(len(shard_tiles(3, r, 8)) for r in range(8)) == 3


def test_tile_outside_the_scene_and_all_nodata(dev):
    """A tile with no pixel inside the scene is all padding (zeros, nodata = 1), like crop_tif/padded_crop
    (src/util/geo_util.py:297-341); nodata pixels of a partly valid scene do not enter the statistics."""
    Hs, Ws, crop = 200, 300, 64
    scene = synth.scene_u16(Hs, Ws, seed=11)
    nodata = np.ones((Hs, Ws), dtype=bool)
    nodata[50:60, 70:90] = False
    boxes = np.array([[-500, -500, -500 + crop, -500 + crop], [Ws + 10, Hs + 10, Ws + 10 + crop, Hs + 10 + crop],
                      [40, 30, 40 + crop, 30 + crop]], dtype=np.int32)
    sc = torch.from_numpy(scene.view(np.int16)).to(dev)
    nd = torch.from_numpy(nodata).to(dev)
    stats = ops.scene_stats(sc, nd)
    out = ops.ingest_tiles(sc, nd, stats, torch.from_numpy(boxes).to(dev), crop, want_u8=True, want_nodata=True)
    assert not out["u8"][:2].any() and out["nodata"][:2].all()
    u8 = glue_ref.tif_image_4band(scene.astype(np.float32), nodata)
    want_u8, want_nd, _ = glue_ref.crop_tif(tuple(int(v) for v in boxes[2]), u8, nodata, None, crop)
    assert np.array_equal(out["u8"][2].cpu().numpy(), want_u8)
    assert np.array_equal(out["nodata"][2].cpu().numpy().astype(bool), want_nd.astype(bool))
    zero = glue_ref.normalize(torch.from_numpy(glue_ref.get_crop_image(np.zeros((crop, crop, 3), np.uint8), 448))[None])[0]
    assert torch.equal(out["image"][0].cpu(), zero)


def test_accumulator_ignores_a_crop_outside_the_canvas(dev):
    """src/predict.py:151-153: an update whose destination is empty is skipped (the reference logs 'Invalid crop!')."""
    acc = Accumulator((100, 120), None, device=dev)
    one_hot = np.eye(4, dtype=np.uint8)[np.full((32, 32), 2)]
    acc.update("d", (500, 500, 532, 532), one_hot)
    assert not acc.current_pred_counter.any()
    acc.update("d", (110, 90, 142, 122), one_hot)  # 10 x 10 pixels inside
    assert int((acc.prediction() == 2).sum()) == 100


def test_c_abi_argument_errors(dev, small_model):
    """-1000 + a message for bad arguments, no launch, no fallback."""
    L = _lib.lib()
    z = torch.zeros((1, 3, 448, 448), device=dev)
    pred = torch.empty((1, 3, 896, 448), device=dev)
    ws = torch.empty(1024, dtype=torch.uint8, device=dev)
    base = (ws.data_ptr() + 255) // 256 * 256
    rc = L.bseg_forward(small_model._handle, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), 1, 0, 0, C.c_void_p(base),
                        C.c_size_t(512), _lib.ptr(pred), _lib.stream_ptr())
    assert rc == -1000 and b"workspace too small" in L.bseg_last_error()
    rc = L.bseg_forward(small_model._handle, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), 0, 0, 0, C.c_void_p(base),
                        C.c_size_t(512), _lib.ptr(pred), _lib.stream_ptr())
    assert rc == -1000
    rc = L.bseg_forward(small_model._handle, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), 3, 0, 2, C.c_void_p(base),
                        C.c_size_t(512), _lib.ptr(pred), _lib.stream_ptr())
    assert rc == -1000 and b"multiple of ensemble_prompts" in L.bseg_last_error()
    rc = L.bseg_forward_f32(small_model._handle, _lib.ptr(z), _lib.ptr(z), _lib.ptr(z), 1, 0, 0, C.c_void_p(base),
                            C.c_size_t(512), _lib.ptr(pred), _lib.stream_ptr())
    assert rc == -1000 and b"bseg_enable_fp32" in L.bseg_last_error()
    with pytest.raises(_lib.BsegError):
        ops.scene_stats(torch.zeros((4, 8, 8), dtype=torch.int16), torch.zeros((8, 8), dtype=torch.bool))  # CPU tensors


def test_cuda_graph_replay_is_bit_identical_to_plain_launches(dev):
    """Small batches go through staging buffers and a CUDA graph of the whole forward (bseg_set_graph_batch_limit): the
    first call runs eagerly, the second captures, later ones replay.  Same kernels in the same order: the results must be
    bit-identical to a module that never uses graphs -- for both embedding types, the feature ensemble and the
    query-half-only decoder, and with inputs that change between calls."""
    from oracle.seggpt_ref import make_reference_model

    hf = make_reference_model(seed=3, stress=True, num_layers=5, merge_index=1, intermediate=(1, 2, 3, 4))
    plain = SegGptB200.from_hf(hf, device=dev, graph_batch=0)
    graphed = SegGptB200.from_hf(hf, device=dev, graph_batch=4)
    cases = [dict(embedding_type="instance"), dict(embedding_type="semantic"),
             dict(feature_ensemble=True), dict(query_half_only=True)]
    with torch.no_grad():
        for kw in cases:
            for call in range(4):  # eager, capture, replay, replay -- new inputs every time
                px, ppx, pm = (t.to(dev) for t in synth.model_inputs(batch=2, seed=400 + call))
                want = plain(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, **kw).pred_masks
                got = graphed(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, **kw).pred_masks
                assert torch.equal(got, want), (kw, call)
        # a batch above the limit takes the plain path
        px, ppx, pm = (t.to(dev) for t in synth.model_inputs(batch=5, seed=9))
        assert torch.equal(graphed(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm).pred_masks,
                           plain(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm).pred_masks)


def test_fused_residual_layernorm_levels(dev, small_model):
    """bseg_gemm_set_fused_ln: with level 1 (lin2 also writes the next layer's norm1) and level 2 (proj also writes
    norm2) the forward must stay within bf16-LayerNorm rounding of level 0 (the statistics are merged in another order:
    single-ulp differences of a few LayerNorm outputs), reproducible run to run, batch-independent, and usable under
    CUDA-graph replay (the exchange buffer is re-initialised inside the captured forward)."""
    L = _lib.lib()
    px, ppx, pm = synth.model_inputs(batch=3, seed=9)
    args = dict(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev))
    prev = L.bseg_gemm_set_fused_ln(0)
    try:
        with torch.no_grad():
            base = small_model(**args).pred_masks.clone()
            for level in (1, 2):
                assert L.bseg_gemm_set_fused_ln(level) == (0 if level == 1 else 1)
                got = small_model(**args).pred_masks.clone()
                again = small_model(**args).pred_masks.clone()   # (second call with the same arguments: graph capture)
                third = small_model(**args).pred_masks.clone()   # (replay)
                assert torch.equal(got, again) and torch.equal(got, third)
                rel = ((got - base).norm() / base.norm()).item()
                print(f"[fused LN level {level}] rel-L2 vs unfused = {rel:.3e}")
                assert rel < 5e-3  # (the bf16 path itself sits 5e-3 .. 7e-3 from fp32 on this stress-initialised model)
                one = small_model(pixel_values=px[2:3].to(dev), prompt_pixel_values=ppx[2:3].to(dev),
                                  prompt_masks=pm[2:3].to(dev)).pred_masks
                assert torch.equal(one[0], got[2])
    finally:
        L.bseg_gemm_set_fused_ln(prev)


def test_programmatic_dependent_launch_is_invisible(dev, small_model):
    """bseg_set_pdl: launching the forward-path kernels with programmatic stream serialization (each calls
    griddepcontrol.wait after its prologue) must not change a single bit -- eager, graph-captured and replayed, batch 1
    and a ragged multi-launch batch -- and the training forward / backward must be unaffected as well."""
    L = _lib.lib()
    prev = L.bseg_set_pdl(0)
    try:
        outs = {}
        for on in (0, 1):
            L.bseg_set_pdl(on)
            res = []
            for batch in (1, 5):
                px, ppx, pm = synth.model_inputs(batch=batch, seed=40 + batch)
                args = dict(pixel_values=px.to(dev), prompt_pixel_values=ppx.to(dev), prompt_masks=pm.to(dev))
                with torch.no_grad():
                    a = small_model(**args).pred_masks.clone()
                    b = small_model(**args).pred_masks.clone()   # capture
                    c = small_model(**args).pred_masks.clone()   # replay
                assert torch.equal(a, b) and torch.equal(a, c)
                res.append(a)
            px, ppx, pm = synth.model_inputs(batch=2, seed=50)
            p = ppx.to(dev).requires_grad_(True)
            out = small_model(pixel_values=px.to(dev), prompt_pixel_values=p, prompt_masks=pm.to(dev))
            d = torch.zeros_like(out.pred_masks)
            d[:, :, 448:] = torch.randn(d[:, :, 448:].shape, generator=torch.Generator().manual_seed(3)).to(dev)
            out.pred_masks.backward(d)
            torch.cuda.synchronize()
            res += [out.pred_masks.detach().clone(), p.grad.clone()]
            outs[on] = res
        for x, y in zip(outs[0], outs[1]):
            assert torch.equal(x, y)
    finally:
        L.bseg_set_pdl(prev)
