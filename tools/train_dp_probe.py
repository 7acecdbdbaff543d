"""Where does the data-parallel train step lose time against the single-GPU step?  torchrun --nproc-per-node 2.
Times the bench's train step (batch 32 per GPU) in four variants on every rank:
  full        begin_step (index all-gather on a side stream) + backward into the persistent buffer + AVG all-reduce
  no-reduce   the same without the all-reduce (finish_step only reads the gathered indices)
  no-exchange exchange disabled altogether (the world-size-1 code path, ranks independent)
  barrier     no-exchange plus a dist.barrier() per step (cost of keeping the ranks in lock step)"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

from beach_seg_b200 import ops, synth
from beach_seg_b200.config import BeachSegConfig
from beach_seg_b200.model import PromptModel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, P, CROP = 32, 32, 512
torch.manual_seed(0)
from beach_seg_b200.ml_util import load_model

model = load_model("random-init:0", device=dev, max_batch=64)
scene_np = synth.scene_u16(CROP * 8, CROP * 8, seed=1000 + rank)
scene = torch.from_numpy(scene_np.view(np.int16)).to(dev)
nodata = torch.zeros(scene_np.shape[1:], dtype=torch.bool, device=dev)
boxes = torch.from_numpy(synth.tile_boxes(64, CROP, CROP * 8)).to(dev)[:B]
stats = ops.scene_stats(scene, nodata)
conf = BeachSegConfig(checkpoint="random-init:0", batch_size=B, world_size=world, epochs=1)
pm = PromptModel(conf, device=dev, model=model)
model.check_grad_support = False
img01, cls = synth.smooth_image(P, 5000), synth.blocky_mask(P, 5001)


class DM:
    prompt_imgs = [{"image": img01[i], "mask": cls[i][None], "crop_idx": i} for i in range(P)]


pm.create_trainable_params(DM)
pm.g.manual_seed(conf.seed + rank)
opt = pm.configure_optimizers()["optimizer"]
labels = synth.blocky_mask(B, 4000 + rank)[:, None].to(dev)
mode = {"v": "full"}
real_exchange = pm._exchange


def step():
    tiles = ops.ingest_tiles(scene, nodata, stats, boxes, CROP)
    loss = pm.training_step({"image": tiles["image"], "mask": labels}, 0)
    loss.backward()
    if mode["v"] == "full":
        pm.sync_prompt_grads()
    elif mode["v"] == "no-reduce":
        ex = real_exchange()
        ex.event.synchronize()
        used = set(ex._host.tolist())
        for i, p in enumerate(ex.params):
            p.grad = ex.views[i] if i in used else None
    elif mode["v"] == "barrier":
        dist.barrier()
    opt.step()
    opt.zero_grad(set_to_none=True)


def timed(n):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for v in ["full", "no-reduce", "no-exchange", "barrier", "full"]:
    mode["v"] = v
    pm._exchange = real_exchange if v in ("full", "no-reduce") else (lambda: None)
    for _ in range(2):
        step()
    ms = timed(5)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    allt = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        print(f"{world} GPUs, {v:12s}: " + "  ".join(f"{float(x):7.2f}" for x in allt) + " ms/iter per rank", flush=True)
dist.destroy_process_group()
