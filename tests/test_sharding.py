"""CPU: host logic of the multi-GPU path (tile sharding, canvas merge) with world_size-2 gloo, and the host mirrors of
the reference's palette / config helpers."""
import dataclasses
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from beach_seg_b200 import synth
from beach_seg_b200.predict import create_palette, shard_tiles
from oracle import glue_ref


def test_shard_tiles_partitions_exactly():
    for n in (0, 1, 7, 64, 162, 163):
        for world in (1, 2, 4, 8):
            parts = [list(shard_tiles(n, r, world)) for r in range(world)]
            flat = [i for p in parts for i in p]
            assert flat == list(range(n))                      # contiguous, ordered, no overlap, nothing lost
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= max(1, (n + world - 1) // world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Hs, Ws, crop = 600, 1000, 256
    boxes = synth.sliding_boxes(Hs, Ws, crop, 192)
    rng = np.random.default_rng(0)
    cls = rng.integers(0, 4, size=(len(boxes), crop, crop)).astype(np.uint8)
    mine = shard_tiles(len(boxes), rank, world)
    acc = glue_ref.AccumulatorRef((Hs, Ws))           # stands in for the per-rank device canvas (same byte layout)
    for i in mine:
        acc.update(tuple(int(v) for v in boxes[i]), np.eye(4, dtype=np.uint8)[cls[i]])
    canvas = torch.from_numpy(acc.counter.view(np.uint32).reshape(Hs, Ws).astype(np.int64))
    dist.reduce(canvas, dst=0, op=dist.ReduceOp.SUM)  # the one collective of a shared scene: add the u32 canvases
    if rank == 0:
        np.save(Path(out_dir) / "merged.npy", canvas.numpy().astype(np.uint32))
    dist.destroy_process_group()


def test_two_rank_canvas_merge_equals_single_process(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    merged = np.load(tmp_path / "merged.npy").view(np.uint8).reshape(600, 1000, 4)
    Hs, Ws, crop = 600, 1000, 256
    boxes = synth.sliding_boxes(Hs, Ws, crop, 192)
    cls = np.random.default_rng(0).integers(0, 4, size=(len(boxes), crop, crop)).astype(np.uint8)
    acc = glue_ref.AccumulatorRef((Hs, Ws))
    for b, c in zip(boxes, cls):
        acc.update(tuple(int(v) for v in b), np.eye(4, dtype=np.uint8)[c])
    assert np.array_equal(merged, acc.counter)


def _grad_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from beach_seg_b200.model import allreduce_prompt_grads

    # 4 prompts; rank 0 touched prompts 0 and 1, rank 1 touched prompts 1 and 3; prompt 2 is unused everywhere
    params = [torch.nn.Parameter(torch.zeros(3, 8, 8)) for _ in range(4)]
    touched = [(0, 1), (1, 3)][rank]
    for i in touched:
        params[i].grad = torch.full((3, 8, 8), float(10 * rank + i + 1))
    allreduce_prompt_grads(params)
    if rank == 0:
        torch.save([None if p.grad is None else p.grad.clone() for p in params], Path(out_dir) / "grads.pt")
    dist.destroy_process_group()


def test_two_rank_prompt_gradient_allreduce(tmp_path):
    """The train path's one collective: dense mean all-reduce, untouched prompts keep grad None (AdamW skips them)."""
    mp.spawn(_grad_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = torch.load(tmp_path / "grads.pt")
    assert g[2] is None
    assert torch.equal(g[0], torch.full((3, 8, 8), 1.0 / 2))            # rank 0 only: (0*10 + 0 + 1) / world
    assert torch.equal(g[1], torch.full((3, 8, 8), (2.0 + 12.0) / 2))   # both ranks
    assert torch.equal(g[3], torch.full((3, 8, 8), 14.0 / 2))           # rank 1 only


def _exchange_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from beach_seg_b200.model import PromptGradExchange

    # 4 prompts; rank 0 drew prompts (0, 1, 1), rank 1 drew (1, 3, 3); prompt 2 is unused everywhere
    params = [torch.nn.Parameter(torch.zeros(3, 8, 8)) for _ in range(4)]
    ex = PromptGradExchange(params)
    out = []
    for step in range(2):  # the buffer is reused: the second step must not see the first one's gradients
        drawn = torch.tensor([(0, 1, 1), (1, 3, 3)][rank])
        ex.begin_step(drawn)
        loss = sum(params[int(i)].sum() * float(10 * rank + int(i) + 1 + step) for i in drawn)
        loss.backward()  # autograd accumulates in place into the rows of the persistent buffer
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, ex.views))
        ex.finish_step()
        out.append([None if p.grad is None else p.grad.clone() for p in params])
        assert all(p.grad is None or p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, ex.views))
    hashes = [torch.zeros(4, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(hashes, torch.stack([torch.zeros((), dtype=torch.float64) if g is None else g.double().sum()
                                         for g in out[-1]]))
    assert all(torch.equal(h, hashes[0]) for h in hashes)  # every rank ends the step with the same gradients
    if rank == 0:
        torch.save(out, Path(out_dir) / "grads.pt")
    dist.destroy_process_group()


def test_two_rank_persistent_prompt_gradient_exchange(tmp_path):
    """PromptGradExchange: p.grad are rows of ONE persistent buffer, one average all-reduce, prompts no rank drew keep
    grad None, no stale gradient survives into the next step."""
    mp.spawn(_exchange_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for step, g in enumerate(torch.load(tmp_path / "grads.pt")):
        assert g[2] is None
        assert torch.equal(g[0], torch.full((3, 8, 8), (1.0 + step) / 2))                          # rank 0, once
        assert torch.equal(g[1], torch.full((3, 8, 8), (2 * (2.0 + step) + (12.0 + step)) / 2))    # rank 0 twice, rank 1 once
        assert torch.equal(g[3], torch.full((3, 8, 8), 2 * (14.0 + step) / 2))                     # rank 1, twice


def test_create_palette_matches_reference_rng_stream():
    """Drop-in: same torch.manual_seed -> same random palette as the reference's CPU path (src/model.py:215-231)."""
    torch.manual_seed(42)
    pal, paln = create_palette(4, 3, True, "cpu")
    torch.manual_seed(42)
    rpal, rpaln = glue_ref.create_palette(4, 3, train=True)
    assert torch.equal(pal, rpal) and torch.equal(paln, rpaln)
    pal, paln = create_palette(4, 2, False, "cpu")
    rpal, rpaln = glue_ref.create_palette(4, 2, train=False)
    assert torch.equal(pal.float(), rpal.float()) and torch.equal(paln, rpaln)


def test_config_mirrors_reference_defaults():
    ref_root = Path("/root/reference")
    if not ref_root.exists():
        pytest.skip("reference tree not present on this box")
    sys.path.insert(0, str(ref_root))
    from src.config import BeachSegConfig as RefConfig  # imports cleanly (PIL only)

    from beach_seg_b200.config import BeachSegConfig

    ref = {f.name: f.default for f in dataclasses.fields(RefConfig)}
    ours = {f.name: f.default for f in dataclasses.fields(BeachSegConfig)}
    assert set(ref) == set(ours)
    for k, v in ref.items():
        if k == "resample":
            assert ours[k] == v.name
        elif k in ("data", "model_training_root"):  # the reference's defaults are the author's private directories
            assert isinstance(ours[k], Path)
        else:
            assert ours[k] == v, k


def test_shard_rows_cover_the_tiles_and_the_scene():
    """predict.shard_rows: the rows a rank uploads contain every row its tiles read, and the statistics rows of all
    ranks partition the scene (src/util/geo_util.py:459-464 needs scene-global min / max)."""
    import numpy as np

    from beach_seg_b200 import synth
    from beach_seg_b200.predict import shard_rows, shard_tiles

    Hs, Ws = 4000, 8000
    boxes = synth.sliding_boxes(Hs, Ws, 512, 448)
    for world in (1, 2, 3, 8, 200):
        covered = np.zeros(Hs, dtype=np.int32)
        for r in range(world):
            (t0, t1), (s0, s1) = shard_rows(boxes, Hs, r, world)
            covered[s0:s1] += 1
            for i in shard_tiles(len(boxes), r, world):
                y0, y1 = max(int(boxes[i][1]), 0), min(int(boxes[i][3]), Hs)
                assert t0 <= y0 and y1 <= t1
            if len(shard_tiles(len(boxes), r, world)) == 0:
                assert (t0, t1) == (0, 0)
        assert (covered == 1).all()
