"""Small-batch forward latency with and without programmatic dependent launch (bseg_set_pdl)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib, synth
from beach_seg_b200.ml_util import load_model

dev = torch.device("cuda:0")
L = _lib.lib()
model = load_model("random-init:0", device=dev, max_batch=64, graph_batch=16)
for B in (1, 2, 4, 16, 64):
    px, ppx, pm = (t.to(dev) for t in synth.model_inputs(batch=B, seed=1))
    res = {}
    for rnd in range(3):
        for on in (0, 1):
            L.bseg_set_pdl(on)
            with torch.no_grad():
                for _ in range(4):
                    model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm)
                torch.cuda.synchronize()
                reps = 30 if B < 64 else 6
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    model(pixel_values=px, prompt_pixel_values=ppx, prompt_masks=pm)
                e1.record()
                torch.cuda.synchronize()
            res.setdefault(on, []).append(e0.elapsed_time(e1) / reps)
    print(f"B={B:2d}: " + "  ".join(f"pdl {k}: {min(v):8.3f} ms/call" for k, v in res.items()), flush=True)
L.bseg_set_pdl(1)
