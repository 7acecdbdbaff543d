"""Per-step device time of the host-buffer pipeline (predict.HostScenePipeline) next to the device-resident step, per
rank.  torchrun --nproc-per-node N (or plain python for one GPU)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
import torch.distributed as dist

from beach_seg_b200 import ops, synth
from beach_seg_b200.ml_util import load_model
from beach_seg_b200.predict import HostScenePipeline, TilePredictor, create_palette

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
CROP, N = 512, 64
model = load_model("random-init:0", device=dev, max_batch=N)
predictor = TilePredictor(model, CROP)
scene_np = synth.scene_u16(CROP * 8, CROP * 8, seed=1000 + rank)
scene_host = torch.from_numpy(scene_np.view(np.int16)).pin_memory()
scene = scene_host.to(dev)
nodata = torch.zeros(scene_np.shape[1:], dtype=torch.bool, device=dev)
boxes = torch.from_numpy(synth.tile_boxes(N, CROP, CROP * 8)).to(dev)
stats = ops.scene_stats(scene, nodata)
prompt_images = synth.normalize(synth.smooth_image(N, 2000 + rank)).to(dev)
prompt_cls = synth.blocky_mask(N, 3000 + rank).to(dev)
torch.manual_seed(42)
palette = create_palette(4, N, True, dev)
canvas = torch.zeros(scene_np.shape[1:], dtype=torch.int32, device=dev)
pipe = HostScenePipeline(predictor, scene_host.shape, N, CROP)


def step_device():
    cls = predictor.predict_tiles(scene, nodata, stats, boxes, prompt_images, prompt_cls, palette)
    ops.vote_accumulate(canvas, cls, boxes, overlapping=False)


def step_e2e():
    pipe.step(scene_host, nodata, stats, boxes, prompt_images, prompt_cls, palette, canvas)


def run(fn, n, tag):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.current_stream().wait_stream(pipe.copy_stream)
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print(f"rank {rank} {tag:8s}: total {ev[0].elapsed_time(end) / n:7.2f} ms/step; steps " +
          " ".join(f"{p:6.1f}" for p in per), flush=True)


for _ in range(3):
    step_device()
run(step_device, 6, "device")
for _ in range(3):
    step_e2e()
pipe.drain()
run(step_e2e, 6, "e2e")
run(step_device, 6, "device")
run(step_e2e, 6, "e2e")
if world > 1:
    dist.destroy_process_group()
