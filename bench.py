#!/usr/bin/env python
"""Benchmark of the beach_seg hot path on B200.

Workload (BASELINE.json configs[1]): batched tile inference, 64 tiles of 512x512 4-band uint16 per step per GPU:
ingest (composite + PIL-bicubic 512->448 + normalise) -> prompt colourise -> SegGPT ViT-L forward (random init,
seed 0) -> palette decode (+nearest resize back to 512) -> vote stitch.  metric = tiles/s (whole job, all GPUs).

    python bench.py --gpus N --steps K --warmup W            # our arm (under torchrun for N > 1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path on the host cores

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for how every field is produced.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

TILES_PER_STEP = 64
TRAIN_BATCH = 32   # BASELINE.json configs[3]: train step, batch 32 of 512^2 tiles per GPU
TRAIN_PROMPTS = 32
TRAIN_FLOP_PER_TILE = 3.456e12  # forward + dgrad-only backward (no wgrad: the backbone is frozen)
CROP = 512
FWD_FLOP_PER_TILE = 1.5897e12  # BASELINE.md §3 / SURVEY §8(d): algorithmic forward FLOPs per 448-path tile
CATS = ("gemm", "attention", "layernorm", "decoder_head", "ingest", "decode", "vote", "elementwise", "loss")


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tensor": d.get("bf16_tflops_sustained", 1393.8), "tensor_burst": d.get("bf16_tflops", 1675.9),
                "hbm": d.get("hbm_gbs", 6441.0), "source": "MEASURED_PEAKS.json"}
    return {"tensor": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# the reference's CPU path (HF SegGPT fp32 eager + restated glue), one tile per step
# ------------------------------------------------------------------------------------------------------------
def cpu_tile_pipeline(n_tiles: int, warmup: int):
    """Times the reference pipeline of src/predict.py:232-262 on the host cores: tif_image + crop + PIL resize +
    normalise -> HF SegGptForImageSegmentation fp32 eager (batch 1, like the reference) -> process_pred_masks ->
    cv2 nearest resize -> Accumulator.update.  Returns (tiles/s, cores, per-tile seconds)."""
    from beach_seg_b200 import synth
    from oracle import glue_ref
    from oracle.seggpt_ref import make_reference_model

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hf = make_reference_model(seed=0, stress=False)
    scene = synth.scene_u16(CROP, CROP * (n_tiles + warmup), seed=1000)
    nodata = np.zeros(scene.shape[1:], dtype=bool)
    ppx = synth.normalize(synth.smooth_image(1, 2000))
    pcls = synth.blocky_mask(1, 3000)
    acc = glue_ref.AccumulatorRef(scene.shape[1:])
    times = []
    u8 = glue_ref.tif_image_4band(scene.astype(np.float32), nodata)  # once per scene, like the reference
    for i in range(n_tiles + warmup):
        t0 = time.perf_counter()
        box = (i * CROP, 0, (i + 1) * CROP, CROP)
        ci, _, _ = glue_ref.crop_tif(box, u8, nodata, None, CROP)
        px = glue_ref.normalize(torch.from_numpy(glue_ref.get_crop_image(ci, 448))[None])
        pal, pal_norm = glue_ref.create_palette(4, 1, train=True)
        pm = glue_ref.normalize(glue_ref.torch_apply_mask_rgb(pal, pcls))
        with torch.no_grad():
            pred = hf(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm, embedding_type="instance").pred_masks
        cls = glue_ref.process_pred_masks(pred, pal_norm)[0].numpy()
        _, one_hot = glue_ref.predict_tail(cls, CROP)
        acc.update(box, one_hot)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    per = float(np.median(times))
    return 1.0 / per, cores, per


WORKLOAD = ("batched tile inference, 512x512x4 uint16 tiles -> 448 SegGPT ViT-L (random init seed 0), batch 64 per GPU "
            "per step: ingest+colourise+forward+decode+vote")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tps, cores, per = cpu_tile_pipeline(max(args.steps, 1), max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "tiles/sec (512^2 4-band) predict", "value": tps, "unit": "tiles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # our arm's workload; a reference step is a bounded sample of it: one tile (the reference's own batch size,
        # src/predict.py runs batch 1) of the same synthetic 512x512x4 shape through the same model
        "config": {"workload": WORKLOAD, "tiles_per_step_per_gpu": TILES_PER_STEP, "crop": CROP,
                   "sample": "1 tile per step, batch 1 (reference CPU path, src/predict.py loop)", "tiles_per_step": 1},
        "cpu_baseline": {"value": tps, "unit": "tiles/s", "cores": cores, "kind": "reference",
                         "sample": f"{args.steps} tiles, 1 tile/step, HF transformers SegGPT fp32 eager (the "
                                   "reference's own dependency; torch.compile unavailable) + restated glue, median"},
        "e2e": {"value": tps, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# train-step leg of our arm
# ------------------------------------------------------------------------------------------------------------
def train_leg(args, dev, rank, world, model, scene, nodata, stats, boxes, barrier, timed, L, peaks):
    from beach_seg_b200 import ops, synth
    from beach_seg_b200.config import BeachSegConfig
    from beach_seg_b200.model import PromptModel

    conf = BeachSegConfig(checkpoint="random-init:0", batch_size=TRAIN_BATCH, world_size=world, epochs=1)
    pmodel = PromptModel(conf, device=dev, model=model)
    model.check_grad_support = False  # the loss kernel's gradient is zero in the prompt half by construction
    prompt_img01 = synth.smooth_image(TRAIN_PROMPTS, 5000)
    prompt_cls = synth.blocky_mask(TRAIN_PROMPTS, 5001)

    class DM:
        prompt_imgs = [{"image": prompt_img01[i], "mask": prompt_cls[i][None], "crop_idx": i}
                       for i in range(TRAIN_PROMPTS)]

    pmodel.create_trainable_params(DM)
    pmodel.g.manual_seed(conf.seed + rank)  # each rank draws its own prompt indices
    opt = pmodel.configure_optimizers()["optimizer"]
    labels = synth.blocky_mask(TRAIN_BATCH, 4000 + rank)[:, None].to(dev)
    tboxes = boxes[:TRAIN_BATCH]

    def step():
        tiles = ops.ingest_tiles(scene, nodata, stats, tboxes, CROP)          # the step's 32 tiles, 512 -> 448
        loss = pmodel.training_step({"image": tiles["image"], "mask": labels}, 0)
        loss.backward()
        pmodel.sync_prompt_grads()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(2):
        loss = step()
    # ---- on-hardware check of the exchange: after sync_prompt_grads every rank holds bit-identical gradients ----
    grads_identical = None
    if world > 1:
        import torch.distributed as dist

        tiles = ops.ingest_tiles(scene, nodata, stats, tboxes, CROP)
        pmodel.training_step({"image": tiles["image"], "mask": labels}, 0).backward()
        pmodel.sync_prompt_grads()
        h = torch.stack([p.grad.view(torch.int32).to(torch.int64).sum() if p.grad is not None
                         else torch.full((), -1, dtype=torch.int64, device=dev) for p in pmodel.prompt_params_list])
        allh = [torch.zeros_like(h) for _ in range(world)]
        dist.all_gather(allh, h)
        grads_identical = all(bool(torch.equal(a, allh[0])) for a in allh)
        if not grads_identical:
            raise SystemExit("train: prompt gradients differ across ranks after sync_prompt_grads")
        opt.zero_grad(set_to_none=True)
    ms, ms_ranks = timed(step, args.train_steps, want_ranks=True)
    # per-category device time of the same steps
    L.bseg_profile_enable(1)
    barrier()
    for _ in range(args.train_steps):
        loss = step()
    torch.cuda.synchronize()
    n = len(CATS)
    pms, pl, pw, pb = (C.c_double * n)(), (C.c_longlong * n)(), (C.c_double * n)(), (C.c_double * n)()
    L.bseg_profile_collect(pms, pl, pw, pb)
    L.bseg_profile_enable(0)
    tot = sum(pms[i] for i in range(n))
    kern = {name: {"ms_per_iter": pms[i] / args.train_steps, "launches_per_iter": pl[i] / args.train_steps,
                   "share": pms[i] / tot if tot else None,
                   "tflops": pw[i] / (pms[i] * 1e-3) / 1e12 if pw[i] and pms[i] else None}
            for i, name in enumerate(CATS) if pl[i]}
    # ---- the same step with the datamodule's train_aug (kornia chain of src/data.py:195-224) in
    # the prompt -> loss autograd chain, and the augmentation kernels timed alone at the same batch ----
    from beach_seg_b200.augment import TrainAug

    aug = TrainAug(conf, generator=torch.Generator().manual_seed(conf.seed + rank))
    pmodel.train_aug = aug
    step()
    ms_aug_step = timed(step, args.train_steps) / args.train_steps
    pmodel.train_aug = pmodel.aug
    img = prompt_img01[torch.arange(TRAIN_BATCH) % TRAIN_PROMPTS].to(dev).requires_grad_(True)
    msk = prompt_cls[torch.arange(TRAIN_BATCH) % TRAIN_PROMPTS].to(dev)
    busy = BeachSegConfig(sharpness_p=1.0, erasing_p=1.0, gauss_p=1.0)    # every optional op on every sample
    aug_busy = TrainAug(busy, generator=torch.Generator().manual_seed(1))
    draw = aug_busy.sample_params(TRAIN_BATCH, 448, 448)
    noise = torch.randn(img.shape, device=dev)
    d_out = torch.randn(img.shape, device=dev)
    outs = {}

    def aug_fwd():
        outs["o"], _ = aug_busy.apply(img, msk, draw, noise=noise)

    def aug_bwd():
        torch.autograd.grad(outs["o"], img, d_out, retain_graph=True)

    aug_fwd(); aug_bwd()
    reps = 20
    ms_aug_fwd = timed(aug_fwd, reps) / reps
    ms_aug_bwd = timed(aug_bwd, reps) / reps
    px = TRAIN_BATCH * 448 * 448
    train_aug = {"ms_per_iter_with_train_aug": ms_aug_step, "batch": TRAIN_BATCH,
                 "fwd_ms": ms_aug_fwd, "bwd_ms": ms_aug_bwd,
                 "fwd_gbs": px * 62 / (ms_aug_fwd * 1e-3) / 1e9, "bwd_gbs": px * 96 / (ms_aug_bwd * 1e-3) / 1e9,
                 "bytes_per_px": {"fwd": 62, "bwd": 96},
                 "note": "fwd/bwd timed through TrainAug.apply / autograd (parameter packing + 2 launches each) with "
                         "sharpen, erase and noise forced on for every sample; bytes = image, colour scratch, noise, "
                         "mask and output traffic of the two passes; frac of HBM peak = gbs / roofline peak of the "
                         "bandwidth kernels"}
    ms_iter = ms / args.train_steps
    tflops = TRAIN_BATCH * TRAIN_FLOP_PER_TILE / (ms_iter * 1e-3) / 1e12
    return {"metric": "train step ms/iter", "ms_per_iter": ms_iter, "batch_per_gpu": TRAIN_BATCH, "n_gpus": world,
            "global_batch": TRAIN_BATCH * world, "tiles_per_s": world * TRAIN_BATCH / (ms_iter * 1e-3),
            "steps": args.train_steps, "loss": float(loss.detach()), "prompts": TRAIN_PROMPTS,
            "loss_semantics": "SegGptLoss as written in the reference (BxB keep-mask broadcast, src/model.py:40-64)",
            "includes": "ingest 512->448, colourise, forward (activations kept), palette decode, loss fwd+bwd, "
                        "backward to the prompts, prompt-grad all-reduce (NCCL, world>1), AdamW step",
            "model_tflops_per_gpu": tflops, "model_frac_of_tensor_peak": tflops / peaks["tensor"],
            "grad_allreduce_bytes": TRAIN_PROMPTS * 3 * 448 * 448 * 4 if world > 1 else 0,
            "grads_identical_on_all_ranks_after_sync": grads_identical,
            "ms_per_iter_per_rank": [m / args.train_steps for m in ms_ranks], "kernels": kern,
            "train_aug": train_aug}


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[2]: full-scene sliding-window predict, 8000x4000 px, 512-px tiles, stride 448 (64-px overlap) = 162
# tiles, sharded over the ranks; vote stitch + one canvas reduce + argmax.  Reference loop: src/predict.py:232-262.
# ------------------------------------------------------------------------------------------------------------
def scene_leg(args, dev, rank, world, model, barrier, timed):
    from beach_seg_b200 import ops, synth
    from beach_seg_b200.predict import (TilePredictor, create_palette, predict_scene, shard_rows, shard_tiles,
                                        upload_scene_rows)

    Hs, Ws, stride = 4000, 8000, 448
    scene_np = synth.scene_u16(Hs, Ws, seed=7)                       # the same scene on every rank
    nodata_np = synth.nodata_wedge(Hs, Ws, 0.05)
    boxes_np = synth.sliding_boxes(Hs, Ws, CROP, stride)
    n = len(boxes_np)
    scene_host = torch.from_numpy(scene_np.view(np.int16)).pin_memory()
    nodata_host = torch.from_numpy(nodata_np).pin_memory()
    scene, nodata = scene_host.to(dev), nodata_host.to(dev)
    boxes = torch.from_numpy(boxes_np).to(dev)
    prompts = synth.normalize(synth.smooth_image(1, 2000)).to(dev).expand(n, -1, -1, -1)
    pcls = synth.blocky_mask(1, 3000).to(dev).expand(n, -1, -1)
    torch.manual_seed(42)
    palette = create_palette(4, n, True, dev)                         # per tile, identical on every rank
    predictor = TilePredictor(model, CROP)
    canvas = torch.zeros((Hs, Ws), dtype=torch.int32, device=dev)
    scene_dev = torch.empty_like(scene)
    nodata_dev = torch.empty_like(nodata)
    pred_host = torch.empty((Hs, Ws), dtype=torch.uint8).pin_memory()
    out = {}

    def run_device():
        out["pred"], _ = predict_scene(predictor, scene, nodata, boxes, prompts, pcls, palette, rank, world,
                                       TILES_PER_STEP, canvas=canvas)

    tile_rows, stat_rows = shard_rows(boxes_np, Hs, rank, world)
    h2d = {}

    def run_e2e():
        # host scene -> device: only the rows this rank's tiles read plus its share of the rows for the scene-global
        # statistics (merged over the ranks with one 4-word all-reduce); predict the shard, reduce the canvases,
        # argmax, class map back to the host on rank 0
        h2d["bytes"] = upload_scene_rows(scene_host, nodata_host, scene_dev, nodata_dev, (tile_rows, stat_rows))
        st = ops.scene_stats_sharded(scene_dev, nodata_dev, stat_rows[0], stat_rows[1])
        pred, _ = predict_scene(predictor, scene_dev, nodata_dev, boxes, prompts, pcls, palette, rank, world,
                                TILES_PER_STEP, stats=st, canvas=canvas)
        if pred is not None:
            pred_host.copy_(pred, non_blocking=True)

    reps = max(args.scene_reps, 1)
    for _ in range(2):
        run_device()
    ms, per_rank = timed(run_device, reps, want_ranks=True)
    scene_dev.fill_(-1)   # rows a rank does not upload must not matter
    nodata_dev.fill_(1)
    run_e2e()
    ms_e2e, _ = timed(run_e2e, reps, want_ranks=True)
    torch.cuda.synchronize()
    e2e_same = None
    if rank == 0:
        e2e_same = bool(torch.equal(pred_host, out["pred"].cpu()))
        if not e2e_same:
            raise SystemExit("config3: the class map of the host-buffer (row-sharded upload) path differs from the "
                             "device-resident path")
    # ---- on-hardware equality: the canvas reduced over `world` ranks == the canvas one rank stitches alone ----
    same = None
    run_device()
    if world > 1:
        reduced = canvas.clone()
        if rank == 0:
            single_pred, single = predict_scene(predictor, scene, nodata, boxes, prompts, pcls, palette, 0, 1,
                                                TILES_PER_STEP)
            same = bool(torch.equal(single, reduced)) and bool(torch.equal(single_pred, out["pred"]))
            if not same:
                raise SystemExit("config3: the canvas reduced over the ranks differs from the single-rank canvas")
    votes_ok = None
    if rank == 0:
        votes = canvas.cpu().numpy().view(np.uint8).reshape(Hs, Ws, 4).sum(axis=2)
        cover = np.zeros((Hs, Ws), dtype=np.int32)
        for x0, y0, x1, y1 in boxes_np:
            cover[max(y0, 0):min(y1, Hs), max(x0, 0):min(x1, Ws)] += 1
        votes_ok = bool(np.array_equal(votes, cover))
        if not votes_ok:
            raise SystemExit("config3: vote totals differ from the tile coverage")
    barrier()
    per_scene, per_scene_e2e = ms / reps, ms_e2e / reps
    return {"workload": f"{Ws}x{Hs} px uint16 4-band scene, {CROP}-px tiles, stride {stride} (64-px overlap): {n} tiles "
                        f"sharded over {world} rank(s) (owner-computes blocks of the row-major tile list), launches of "
                        f"<= {TILES_PER_STEP} tiles, vote stitch, one NCCL sum-reduce of the u32 canvases, argmax",
            "tiles": n, "tiles_per_rank": len(shard_tiles(n, 0, world)), "n_gpus": world,
            "scene_ms": per_scene, "value": n / (per_scene * 1e-3), "unit": "tiles/s",
            "scene_ms_per_rank_min_max": [min(per_rank) / reps, max(per_rank) / reps],
            "e2e": {"scene_ms": per_scene_e2e, "value": n / (per_scene_e2e * 1e-3), "unit": "tiles/s",
                    "h2d_bytes_per_scene_per_rank": h2d["bytes"],
                    "h2d_rows": "rows the rank's tiles read + its 1/N of the rows for the scene statistics",
                    "class_map_equals_device_resident_path": e2e_same,
                    "d2h_bytes_per_scene": int(pred_host.numel())},
            "canvas_reduce_bytes": int(canvas.numel() * 4) if world > 1 else 0,
            "canvas_equals_single_rank_canvas": same, "vote_totals_equal_tile_coverage": votes_ok, "reps": reps}


# ------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: predict_no_prompt.py path on 1024x1024 uint8 RGB crops, n_prompts = 2 with feature_ensemble
# ("batch 16" = 8 tiles x 2 prompts per launch; 16 tiles x 2 = 32 samples also reported).  src/predict_no_prompt.py:270-306.
# ------------------------------------------------------------------------------------------------------------
def noprompt_leg(args, dev, rank, world, model, barrier, timed):
    from beach_seg_b200 import ops, synth
    from beach_seg_b200.ml_util import build_palette
    from beach_seg_b200.predict import NoPromptPredictor

    crop, P = 1024, 2
    res = {"workload": f"predict_no_prompt: {crop}x{crop} uint8 RGB crops -> HF-processor resize to 448, {P} prompts per "
                       "tile with feature_ensemble=True (grouped per tile), mean over prompts, semantic post-process at "
                       "crop size, nodata zeroing, vote", "crop": crop, "prompts_per_tile": P, "n_gpus": world}
    predictor = NoPromptPredictor(model, None, crop)
    pal_u8 = torch.tensor(build_palette(3), dtype=torch.uint8)
    for n in (8, 16):
        g = torch.Generator().manual_seed(900 + rank)
        crops_host = torch.randint(0, 256, (n, crop, crop, 3), generator=g, dtype=torch.uint8).pin_memory()
        crops = crops_host.to(dev)
        nodata = torch.from_numpy(synth.nodata_wedge(crop, crop, 0.05)).to(dev)[None].expand(n, -1, -1).contiguous()
        ppx = synth.normalize(synth.smooth_image(n * P, 7000 + rank)).to(dev)
        pmask = ops.colorize_resize_norm255(synth.blocky_mask(n * P, 7100 + rank).to(dev), pal_u8, 448)
        boxes = torch.from_numpy(synth.tile_boxes(n, crop, crop * 4)).to(dev)
        canvas = torch.zeros((crop * ((n + 3) // 4), crop * 4), dtype=torch.int32, device=dev)
        crops_dev = torch.empty_like(crops)
        cls_host = torch.empty((n, crop, crop), dtype=torch.uint8).pin_memory()

        def step_device():
            cls = predictor.predict_tiles(crops, nodata, ppx, pmask)
            ops.vote_accumulate(canvas, cls, boxes, overlapping=False)

        def step_e2e():
            crops_dev.copy_(crops_host, non_blocking=True)
            cls = predictor.predict_tiles(crops_dev, nodata, ppx, pmask)
            ops.vote_accumulate(canvas, cls, boxes, overlapping=False)
            cls_host.copy_(cls, non_blocking=True)

        for _ in range(3):
            step_device()
        ms, _ = timed(step_device, args.steps, want_ranks=True)
        step_e2e()
        ms_e2e, _ = timed(step_e2e, args.steps, want_ranks=True)
        tps = world * n * args.steps / (ms * 1e-3)
        res[f"tiles{n}_x{P}"] = {"samples_per_launch": n * P, "ms_per_launch": ms / args.steps, "value": tps,
                                 "unit": "tiles/s",
                                 "e2e": {"value": world * n * args.steps / (ms_e2e * 1e-3), "unit": "tiles/s",
                                         "h2d_bytes_per_step": int(crops_host.numel()),
                                         "d2h_bytes_per_step": int(cls_host.numel())},
                                 "model_tflops_per_gpu": tps / world * P * FWD_FLOP_PER_TILE / 1e12}
        del crops, crops_dev, canvas, ppx, pmask
    return res


# ------------------------------------------------------------------------------------------------------------
# native-resolution mode: 512-px tiles WITHOUT the resize to 448, i.e. the backbone at
# SegGptConfig(image_size=(1024, 512)): 64 x 32 tokens, T = 2048; ingest is then a purely HBM-bound per-pixel pass
# ------------------------------------------------------------------------------------------------------------
SETTLE_STEPS = 8  # untimed steps before the first timed region (>= the W the caller asks for)
NATIVE = {512: (32, 2.186e12), 1024: (8, 14.356e12)}  # tile -> (tiles per step, SURVEY section 8(d) forward FLOPs per tile)


def native_leg(args, dev, rank, world, scene, nodata, stats, barrier, timed, L, peaks):
    from beach_seg_b200 import ops, synth
    from beach_seg_b200.ml_util import load_model
    from beach_seg_b200.predict import TilePredictor, create_palette

    res = {}
    for tile, (n, flop) in NATIVE.items():
        model = load_model("random-init:0", device=dev, max_batch=n, image_size=tile)
        predictor = TilePredictor(model, tile)
        prompts = synth.normalize(synth.smooth_image(n, 8000 + rank, size=tile)).to(dev)
        pcls = synth.blocky_mask(n, 8100 + rank, size=tile).to(dev)
        torch.manual_seed(43)
        palette = create_palette(4, n, True, dev)
        canvas = torch.zeros(scene.shape[1:], dtype=torch.int32, device=dev)
        b = torch.from_numpy(synth.tile_boxes(n, tile, scene.shape[2])).to(dev)

        def step():
            cls = predictor.predict_tiles(scene, nodata, stats, b, prompts, pcls, palette)
            ops.vote_accumulate(canvas, cls, b, overlapping=False)

        for _ in range(3):
            step()
        ms, _ = timed(step, args.steps, want_ranks=True)
        L.bseg_profile_enable(1)
        barrier()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        ncat = len(CATS)
        pms, pl, pw, pb = (C.c_double * ncat)(), (C.c_longlong * ncat)(), (C.c_double * ncat)(), (C.c_double * ncat)()
        L.bseg_profile_collect(pms, pl, pw, pb)
        L.bseg_profile_enable(0)
        tot = sum(pms[i] for i in range(ncat))
        kern = {name: {"ms_per_step": pms[i] / args.steps, "share": pms[i] / tot if tot else None,
                       "tflops": pw[i] / (pms[i] * 1e-3) / 1e12 if pw[i] and pms[i] else None,
                       "gbs": pb[i] / (pms[i] * 1e-3) / 1e9 if pb[i] and pms[i] else None}
                for i, name in enumerate(CATS) if pl[i]}
        tps = world * n * args.steps / (ms * 1e-3)
        ing = kern.get("ingest", {})
        res[f"tile{tile}"] = {
            "workload": f"native-resolution mode: {n} tiles of {tile}x{tile}x4 uint16 per GPU per step, backbone at "
                        f"image_size {tile} ({tile // 8}x{tile // 16} tokens, T={2 * (tile // 16) ** 2}, random init seed 0): "
                        "ingest (no resize) + colourise + forward + decode + vote",
            "tiles_per_step_per_gpu": n, "n_gpus": world, "value": tps, "unit": "tiles/s", "ms_per_step": ms / args.steps,
            "flop_per_tile": flop, "model_tflops_per_gpu": tps / world * flop / 1e12,
            "model_frac_of_tensor_peak": tps / world * flop / 1e12 / peaks["tensor"], "kernels": kern,
            "ingest_roofline": ({"bound": "hbm", "achieved": ing["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                                 "frac": ing["gbs"] / peaks["hbm"],
                                 "bytes_per_pixel": "9 in (4 x u16 + nodata) + 12 out (3 x fp32)"}
                                if ing.get("gbs") else None)}
        del model, predictor, prompts, pcls, canvas
        torch.cuda.empty_cache()
    return res

# ------------------------------------------------------------------------------------------------------------
# small-batch latency: the reference's real call pattern is batch 1 (src/data.py:287-293, src/predict.py:234)
# ------------------------------------------------------------------------------------------------------------
def latency_leg(args, dev, model, predictor, scene, nodata, stats, boxes, prompt_images, prompt_cls, palette, timed):
    out = {"what": "ingest + colourise + forward + decode of ONE launch of B tiles (device-resident inputs), averaged "
                   "over back-to-back launches; floor = one pass over the 741 MB of bf16 weights per launch at the "
                   "measured HBM rate, or the launch's FLOPs at the sustained tensor peak, whichever is larger"}
    peaks = measured_peaks()
    for B in (1, 4, 16):
        args_b = (scene, nodata, stats, boxes[:B], prompt_images[:B], prompt_cls[:B], (palette[0][:B], palette[1][:B]))

        def step():
            predictor.predict_tiles(*args_b)

        for _ in range(3):
            step()
        reps = 20
        ms, _ = timed(step, reps, want_ranks=True)
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        host_ms = (time.perf_counter() - t0) * 1e3 / reps   # host enqueue time per launch (no sync inside)
        torch.cuda.synchronize()
        floor = max(0.741e9 / (peaks["hbm"] * 1e9), B * FWD_FLOP_PER_TILE / (peaks["tensor"] * 1e12)) * 1e3
        out[f"batch{B}"] = {"ms_per_launch": ms / reps, "ms_per_tile": ms / reps / B, "host_enqueue_ms": host_ms,
                            "floor_ms_per_launch": floor, "frac_of_floor": floor / (ms / reps)}
        if B == 1:  # programmatic dependent launch (bseg_set_pdl; on for <= 2 tiles by default) off / on, interleaved
            from beach_seg_b200 import _lib
            L = _lib.lib()
            prev = L.bseg_set_pdl(-1)
            ab = {0: [], 1: []}
            for _ in range(3):
                for on in (0, 1):
                    L.bseg_set_pdl(on)
                    for _ in range(3):
                        step()
                    ms_ab, _ = timed(step, reps, want_ranks=True)
                    ab[on].append(ms_ab / reps)
            L.bseg_set_pdl(prev)
            out["batch1"]["pdl_ab"] = {"ms_per_launch_without_pdl": min(ab[0]), "ms_per_launch_with_pdl": min(ab[1]),
                                       "all_without": ab[0], "all_with": ab[1]}
            # the very first timed region of this leg follows the 64-tile legs and carries their clock transient: report
            # the steady value of the default setting (the three interleaved rounds agree to 0.1 %)
            best = min(ab[1 if prev else 0])
            out["batch1"].update({"ms_per_launch_first_region": out["batch1"]["ms_per_launch"], "ms_per_launch": best,
                                  "ms_per_tile": best, "frac_of_floor": floor / best})
    return out


# ------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-tiles", type=int, default=6)
    ap.add_argument("--no-train", action="store_true", help="skip the train-step leg (BASELINE configs[3])")
    ap.add_argument("--train-steps", type=int, default=3)
    ap.add_argument("--no-fast-path", action="store_true", help="skip the query-half-only decoder leg")
    ap.add_argument("--no-scene", action="store_true", help="skip the full-scene leg (BASELINE configs[2])")
    ap.add_argument("--scene-reps", type=int, default=2)
    ap.add_argument("--no-noprompt", action="store_true", help="skip the predict_no_prompt leg (BASELINE configs[4])")
    ap.add_argument("--no-latency", action="store_true", help="skip the batch 1/4/16 latency leg")
    ap.add_argument("--no-native", action="store_true", help="skip the native-resolution legs (512- / 1024-px tiles)")
    ap.add_argument("--no-fp32-check", action="store_true",
                    help="skip the fp32-accuracy-mode leg (bf16 path vs bseg_forward_f32 on the bench inputs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # stdout must carry exactly ONE JSON line: anything libraries print there (NCCL banners, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for our arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)

    from beach_seg_b200 import _lib, ops, synth
    from beach_seg_b200.ml_util import load_model
    from beach_seg_b200.predict import TilePredictor, create_palette

    L = _lib.lib()
    model = load_model("random-init:0", device=dev, max_batch=TILES_PER_STEP)
    predictor = TilePredictor(model, CROP)

    # ---- synthetic inputs: each rank owns its own 64-tile scene stripe (weak scaling, no data-path collective) ----
    side = 8
    scene_np = synth.scene_u16(CROP * side, CROP * side, seed=1000 + rank)
    nodata_np = np.zeros(scene_np.shape[1:], dtype=bool)
    boxes_np = synth.tile_boxes(TILES_PER_STEP, CROP, CROP * side)
    scene_host = torch.from_numpy(scene_np.view(np.int16)).pin_memory()
    scene = scene_host.to(dev)
    nodata = torch.from_numpy(nodata_np).to(dev)
    boxes = torch.from_numpy(boxes_np).to(dev)
    stats = ops.scene_stats(scene, nodata)  # once per scene, like tif_image
    prompt_images = synth.normalize(synth.smooth_image(TILES_PER_STEP, 2000 + rank)).to(dev)
    prompt_cls = synth.blocky_mask(TILES_PER_STEP, 3000 + rank).to(dev)
    torch.manual_seed(42)
    palette = create_palette(4, TILES_PER_STEP, True, dev)
    canvas = torch.zeros(scene_np.shape[1:], dtype=torch.int32, device=dev)
    cls_host = torch.empty((TILES_PER_STEP, CROP, CROP), dtype=torch.uint8)  # (shape only: D2H bytes per step)

    def step_device():
        cls = predictor.predict_tiles(scene, nodata, stats, boxes, prompt_images, prompt_cls, palette)
        ops.vote_accumulate(canvas, cls, boxes, overlapping=False)
        return cls

    from beach_seg_b200.predict import HostScenePipeline

    pipe = HostScenePipeline(predictor, scene_host.shape, TILES_PER_STEP, CROP)

    def step_e2e():
        # public host-buffer API: H2D of this step's 64 uint16 tiles from pinned memory, predict, vote, D2H of the
        # class maps; the copies run on a side stream and overlap the neighbouring steps' compute
        pipe.step(scene_host, nodata, stats, boxes, prompt_images, prompt_cls, palette, canvas)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, want_ranks=False):
        """Device time of `steps` calls, barrier + synchronize on both sides, MAX over ranks (and every rank's own
        time when asked: attributes a straggler to a GPU)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        ranks = [ms]
        if world > 1:
            t = torch.zeros(world, dtype=torch.float64, device=dev)
            t[rank] = ms
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ranks = [float(v) for v in t.tolist()]
            ms = max(ranks)
        barrier()
        return (ms, ranks) if want_ranks else ms

    # W warm-up steps as asked, then untimed "settle" steps up to SETTLE_STEPS in total: the part is power-capped, and
    # for the first ~0.5 s after an idle period it runs 5-8 % above its sustained clock (tools/e2e_probe.py: steps of
    # 109, 116, 116, 122, 123, 120 ms).  Without them `value` carries that transient and the later legs (e2e first of
    # all) do not, which reads as a host-buffer cost that does not exist.
    settle = max(SETTLE_STEPS - args.warmup, 0)
    for _ in range(args.warmup + settle):
        step_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.bseg_launch_count()
    ms, ms_ranks = timed(step_device, args.steps, want_ranks=True)
    launches = int(L.bseg_launch_count() - launches0)
    clocks = sampler.stop() if rank == 0 else None
    value = world * TILES_PER_STEP * args.steps / (ms * 1e-3)

    for _ in range(max(args.warmup, 3)):  # the host-buffer leg gets its own warm-up (both buffer slots, allocator pool)
        step_e2e()
    pipe.drain()

    def timed_e2e(steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_e2e()
        torch.cuda.current_stream().wait_stream(pipe.copy_stream)  # the last download is inside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    ms_e2e = timed_e2e(args.steps)
    e2e_value = world * TILES_PER_STEP * args.steps / (ms_e2e * 1e-3)

    # ---- roofline leg: same steps with per-launch CUDA events on the launching stream ----
    L.bseg_profile_enable(1)
    barrier()
    for _ in range(args.steps):
        step_device()
    torch.cuda.synchronize()
    n = len(CATS)
    pms, pl, pw, pb = (C.c_double * n)(), (C.c_longlong * n)(), (C.c_double * n)(), (C.c_double * n)()
    L.bseg_profile_collect(pms, pl, pw, pb)
    gms, gwk = (C.c_double * 16)(), (C.c_double * 16)()
    L.bseg_profile_collect_gemm(gms, gwk)
    L.bseg_profile_enable(0)
    gemm_modes = {}
    names = {0: "bf16", 1: "lin1_gelu", 2: "proj_f32(ensemble)", 3: "proj_resid", 4: "qkv", 5: "patch_embed",
             6: "dec_embed_pixshuf", 11: "lin2_resid", 14: "dec_embed_pixshuf", 12: "proj_resid_ln", 13: "lin2_resid_ln"}
    for i in range(16):
        if gms[i] > 0:
            gemm_modes[names.get(i, str(i))] = {"ms_per_step": gms[i] / args.steps,
                                                "tflops": gwk[i] / (gms[i] * 1e-3) / 1e12}
    peaks = measured_peaks()
    kernels = {}
    tot_ms = sum(pms[i] for i in range(n))
    for i, name in enumerate(CATS):
        if pl[i] == 0:
            continue
        kernels[name] = {"ms_per_step": pms[i] / args.steps, "launches_per_step": pl[i] / args.steps,
                         "share": pms[i] / tot_ms if tot_ms else None,
                         "tflops": pw[i] / (pms[i] * 1e-3) / 1e12 if pw[i] and pms[i] else None,
                         "gbs": pb[i] / (pms[i] * 1e-3) / 1e9 if pb[i] and pms[i] else None}
    g = CATS.index("gemm")
    gemm_tflops = pw[g] / (pms[g] * 1e-3) / 1e12 if pms[g] else 0.0
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_kernel (all GEMM launches of a step)",
                "achieved": gemm_tflops, "peak": peaks["tensor"], "unit": "TFLOP/s",
                "frac": gemm_tflops / peaks["tensor"], "peak_source": peaks["source"] + " bf16_tflops_sustained",
                "traffic": None, "avg_launch_ms": pms[g] / max(pl[g], 1),
                "flop_per_launch": pw[g] / max(pl[g], 1)}
    tfile = ROOT / "profiles" / "r02_gemm_traffic.json"  # (latest capture; r01_gemm_traffic.json: 2.344 GB per launch)
    if not tfile.exists():
        tfile = ROOT / "profiles" / "r01_gemm_traffic.json"
    if tfile.exists():  # DRAM bytes per launch from the committed `ncu --set full` capture of the layer GEMMs
        t = json.loads(tfile.read_text())
        roofline["traffic"] = t["per_launch_avg_dram_bytes"]
        roofline["traffic_algorithmic_bytes"] = t["per_launch_avg_algorithmic_bytes"]
        roofline["traffic_source"] = t["source"] + " (4 encoder-layer GEMM launches, M=200704)"

    # ---- A/B of the two GEMM kernel variants on the same device-resident step (results are bit-identical) ----
    gemm_variants = {"default": "cta_pairs" if L.bseg_gemm_set_cta_pairs(-1) else "one_cta"}
    cur = L.bseg_gemm_set_cta_pairs(-1)
    for vname, v in (("one_cta", 0), ("cta_pairs", 1)):
        L.bseg_gemm_set_cta_pairs(v)
        step_device()
        ms_v = timed(step_device, args.steps)
        gemm_variants[vname + "_tiles_per_s"] = world * TILES_PER_STEP * args.steps / (ms_v * 1e-3)
    L.bseg_gemm_set_cta_pairs(cur)
    # ---- A/B of the residual + LayerNorm fusion (EPI_RESID_LN) on the same step, interleaved twice so that the
    # power-capped clock drifts hit both sides alike ----
    cur = L.bseg_gemm_set_fused_ln(-1)
    ab = {0: [], 1: [], 2: []}
    for _ in range(2):
        for v in (0, 1, 2):
            L.bseg_gemm_set_fused_ln(v)
            step_device()
            ab[v].append(world * TILES_PER_STEP * args.steps / (timed(step_device, args.steps) * 1e-3))
    L.bseg_gemm_set_fused_ln(cur)
    gemm_variants["fused_residual_layernorm"] = {"default_level": int(cur), "off_tiles_per_s": sum(ab[0]) / 2,
                                                 "lin2_tiles_per_s": sum(ab[1]) / 2,
                                                 "lin2_and_proj_tiles_per_s": sum(ab[2]) / 2}

    # ---- the other kernels against their own bounds (same profiled pass; per-step times) ----
    other = {}
    if "attention" in kernels and clocks and clocks.get("sm_mhz"):
        # one exponential per score element: 16 heads x 1568^2 per sequence and layer; MUFU.EX2 = 16 /clk/SM x 148 SMs
        seqs = TILES_PER_STEP * (2 * (model.merge_index + 1) + (model.num_layers - model.merge_index - 1))
        exps = seqs * 16 * 1568.0 * 1568.0
        peak = 16 * 148 * clocks["sm_mhz"] * 1e6
        ach = exps / (kernels["attention"]["ms_per_step"] * 1e-3)
        other["attention"] = {"bound": "mufu (ex2 at the sampled SM clock)", "achieved": ach / 1e12, "peak": peak / 1e12,
                              "unit": "T exp/s", "frac": ach / peak}
    for name in ("layernorm", "ingest", "decode", "vote"):
        if name in kernels and kernels[name]["gbs"]:
            other[name] = {"bound": "hbm", "achieved": kernels[name]["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                           "frac": kernels[name]["gbs"] / peaks["hbm"]}

    # ---- train-step leg (BASELINE configs[3]): ingest -> PromptModel.training_step -> backward -> prompt-grad
    # all-reduce -> AdamW, batch 32 per GPU (weak scaling), ms/iter = max over ranks ----
    train = None
    if not args.no_train:
        train = train_leg(args, dev, rank, world, model, scene, nodata, stats, boxes, barrier, timed, L, peaks)

    scene3 = None if args.no_scene else scene_leg(args, dev, rank, world, model, barrier, timed)
    noprompt5 = None if args.no_noprompt else noprompt_leg(args, dev, rank, world, model, barrier, timed)
    native = None if args.no_native else native_leg(args, dev, rank, world, scene, nodata, stats, barrier, timed, L,
                                                    peaks)
    latency = None
    if not args.no_latency and world == 1:
        latency = latency_leg(args, dev, model, predictor, scene, nodata, stats, boxes, prompt_images, prompt_cls,
                              palette, timed)

    # ---- optional fast path: decoder on the query half only (bseg_forward_query_half); class maps are bit-identical
    # (checked below).  Reported beside the headline, which keeps the full forward. ----
    fast = None
    if not args.no_fast_path:
        fast_pred = TilePredictor(model, CROP, query_half_only=True)
        fast_pipe = HostScenePipeline(fast_pred, scene_host.shape, TILES_PER_STEP, CROP)
        canvas2 = torch.zeros_like(canvas)

        def step_fast():
            cls = fast_pred.predict_tiles(scene, nodata, stats, boxes, prompt_images, prompt_cls, palette)
            ops.vote_accumulate(canvas2, cls, boxes, overlapping=False)
            return cls

        same = bool(torch.equal(step_fast(), step_device()))
        ms_fast = timed(step_fast, args.steps)
        fast_pipe.step(scene_host, nodata, stats, boxes, prompt_images, prompt_cls, palette, canvas2)
        fast_pipe.drain()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fast_pipe.step(scene_host, nodata, stats, boxes, prompt_images, prompt_cls, palette, canvas2)
        torch.cuda.current_stream().wait_stream(fast_pipe.copy_stream)  # the last download is inside the timed region
        e1.record()
        torch.cuda.synchronize()
        ms_fast_e2e = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_fast_e2e, op=dist.ReduceOp.MAX)
        barrier()
        fast = {"what": "decoder_embed + conv head on the query half only (bseg_forward_query_half); the reference's "
                        "predict path never reads pred_masks[:, :, :448] (src/model.py:158-160)",
                "value": world * TILES_PER_STEP * args.steps / (ms_fast * 1e-3),
                "e2e": world * TILES_PER_STEP * args.steps / (ms_fast_e2e.item() * 1e-3), "unit": "tiles/s",
                "class_maps_identical_to_full_forward": same}

    # ---- accuracy leg: the first 8 tiles of the step through the fp32 mode (bseg_forward_f32) and through the bf16
    # path, same inputs: logit deviation and class-map disagreements (north_star: "with that count reported") ----
    fp32_mode = None
    if rank == 0 and world == 1 and not args.no_fp32_check:
        nb = 8
        model32 = load_model("random-init:0", device=dev, max_batch=nb, precision="fp32")
        tiles = ops.ingest_tiles(scene, nodata, stats, boxes[:nb], CROP)["image"]
        pcol = ops.colorize_norm(prompt_cls[:nb], palette[0][:nb])
        kw = dict(pixel_values=tiles, prompt_pixel_values=prompt_images[:nb], prompt_masks=pcol,
                  embedding_type="instance")
        with torch.no_grad():
            p32 = model32(**kw).pred_masks  # warm-up + result
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p32 = model32(**kw).pred_masks
            e1.record()
            torch.cuda.synchronize()
            p16 = model(**kw).pred_masks
        c32 = ops.decode_palette(p32, palette[1][:nb], out_size=CROP, dtype=torch.uint8)
        c16 = ops.decode_palette(p16, palette[1][:nb], out_size=CROP, dtype=torch.uint8)
        q32, q16 = p32[:, :, 448:], p16[:, :, 448:]
        fp32_mode = {"tiles": nb, "tiles_per_s": nb / (e0.elapsed_time(e1) * 1e-3),
                     "bf16_vs_fp32_logits_rel_l2": ((q16 - q32).norm() / q32.norm()).item(),
                     "bf16_vs_fp32_max_abs_over_max": ((q16 - q32).abs().max() / q32.abs().max()).item(),
                     "class_map_disagreements": int((c32 != c16).sum().item()), "class_map_pixels": int(c32.numel()),
                     "note": "fp32 mode = every operand/accumulator IEEE fp32 on the CUDA cores; its own deviation from "
                             "the HF fp32 CPU forward is 2e-6 rel-L2 (tests/test_gpu_fp32_mode.py)"}
        # the same comparison for the train step's gradient (bseg_forward_train_f32 + bseg_backward_to_prompt_f32 against
        # the tensor-core path), 2 tiles, d(pred) = the reference loss's gradient on the bf16 prediction
        ng = 2
        labels = ops.colorize_norm(prompt_cls[ng:2 * ng], palette[0][ng:2 * ng])
        yes = torch.ones((ng, 448, 448), dtype=torch.bool, device=dev)
        kwg = dict(pixel_values=tiles[:ng], prompt_masks=pcol[:ng], embedding_type="instance")
        grads, t_step = {}, {}
        d_pred = None
        for name, m in (("bf16", model), ("fp32", model32)):
            ppx = prompt_images[:ng].clone().requires_grad_(True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            pred = m(prompt_pixel_values=ppx, **kwg).pred_masks
            if d_pred is None:
                _, d_pred = ops.smooth_l1_loss(pred.detach(), labels, yes, 0.01, per_sample=True, want_grad=True)
            pred.backward(d_pred)
            e1.record()
            torch.cuda.synchronize()
            grads[name], t_step[name] = ppx.grad, e0.elapsed_time(e1)
        fp32_mode["train_step"] = {
            "tiles": ng, "fp32_ms": t_step["fp32"],
            "bf16_vs_fp32_prompt_grad_rel_l2": ((grads["bf16"] - grads["fp32"]).norm() / grads["fp32"].norm()).item(),
            "cosine": torch.nn.functional.cosine_similarity(grads["bf16"].flatten(), grads["fp32"].flatten(), dim=0).item(),
            "note": "the fp32 train step is within 5e-6 rel-L2 of torch autograd through the HF module "
                    "(tests/test_gpu_fp32_mode.py)"}
        del model32, p32, p16, grads

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        tps, cores, per = cpu_tile_pipeline(args.cpu_tiles, 1)
        cpu_baseline = {"value": tps, "unit": "tiles/s", "cores": cores, "kind": "reference",
                        "sample": f"{args.cpu_tiles} tiles (1 warm-up), batch 1, same synthetic tile shape: HF "
                                  "transformers SegGPT fp32 eager (the reference's own dependency; torch.compile "
                                  "unavailable in this image) + restated glue, median per tile"}

    if rank == 0:
        line = {
            "metric": "tiles/sec (512^2 4-band) predict", "value": value, "unit": "tiles/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "tiles_per_step_per_gpu": TILES_PER_STEP, "crop": CROP, "parallelism": f"dp{world} (tile shards, "
                       "no data-path collective)", "l2": "per-step activations (>8 GB) exceed the 126 MB L2",
                       "untimed_steps_before_timing": args.warmup + settle},
            "e2e": {"value": e2e_value, "unit": "tiles/s", "h2d_bytes_per_step": int(scene_host.numel() * 2),
                    "d2h_bytes_per_step": int(cls_host.numel()), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_other_kernels": other,
            "cpu_baseline": cpu_baseline,
            "kernels": kernels, "gemm_modes": gemm_modes, "gemm_variants": gemm_variants, "train": train, "fp32_mode": fp32_mode,
            "query_half_fast_path": fast,
            "config3_scene": scene3, "config5_no_prompt": noprompt5, "small_batch_latency": latency,
            "native_resolution": native,
            "ms_per_step_per_rank": [m / args.steps for m in ms_ranks],
            "model_tflops": value / world * FWD_FLOP_PER_TILE / 1e12,
            "model_frac_of_tensor_peak": value / world * FWD_FLOP_PER_TILE / 1e12 / peaks["tensor"],
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
