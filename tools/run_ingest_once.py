"""Runs the tile-ingest kernel on the bench workload (64 tiles of 512x512x4 uint16 -> 448, fp32 NCHW out) for timing and
ncu.  usage: run_ingest_once.py [iters] [crop]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from beach_seg_b200 import ops, synth

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device("cuda:0")
side = 8
scene_np = synth.scene_u16(crop * side, crop * side, seed=1000)
scene = torch.from_numpy(scene_np.view(np.int16)).to(dev)
nodata = torch.zeros(scene_np.shape[1:], dtype=torch.bool, device=dev)
boxes = torch.from_numpy(synth.tile_boxes(64, crop, crop * side)).to(dev)
stats = ops.scene_stats(scene, nodata)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.ingest_tiles(scene, nodata, stats, boxes, crop)
torch.cuda.synchronize()
ms = []
for _ in range(iters):
    flush.zero_()  # evict the scene from L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.ingest_tiles(scene, nodata, stats, boxes, crop)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
ms = sorted(ms)
med = ms[len(ms) // 2]
nbytes = 64 * (9 * crop * crop + 3 * 448 * 448 * 4)
print(f"ingest crop={crop}: median {med * 1e3:.1f} us  min {ms[0] * 1e3:.1f} us  algorithmic {nbytes / 1e6:.1f} MB -> "
      f"{nbytes / med / 1e6:.0f} GB/s")
