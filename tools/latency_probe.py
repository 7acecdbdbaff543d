"""Small-batch forward latency with and without the CUDA-graph path (device time per call with CUDA events, host time per
call without synchronising)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import synth
from beach_seg_b200.ml_util import load_model

dev = torch.device("cuda:0")
for graph_batch in (0, 16):
    model = load_model("random-init:0", device=dev, max_batch=64, graph_batch=graph_batch)
    for B in (1, 2, 4, 8, 16):
        px, ppx, pm = (t.to(dev) for t in synth.model_inputs(batch=B, seed=1))
        with torch.no_grad():
            for _ in range(4):
                model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm)
            torch.cuda.synchronize()
            reps = 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm)
            e1.record()
            torch.cuda.synchronize()
            dev_ms = e0.elapsed_time(e1) / reps
            t0 = time.perf_counter()
            for _ in range(reps):
                model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm)
            host_ms = (time.perf_counter() - t0) * 1e3 / reps
            torch.cuda.synchronize()
        print(f"graph_batch={graph_batch:2d} B={B:2d}: device {dev_ms:7.3f} ms/call ({dev_ms / B:6.3f} ms/tile)  host enqueue "
              f"{host_ms:6.3f} ms/call")
    del model
