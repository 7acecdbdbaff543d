"""Is one 64-tile forward faster or slower than 2 x 32 / 4 x 16 / 8 x 8 (activations of a smaller launch partly stay in
the 126 MB L2 between the kernels of a layer)?  Times model(...) on device-resident inputs."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200.ml_util import load_model

dev = torch.device("cuda:0")
model = load_model("random-init:0", device=dev, max_batch=64, graph_batch=0)
g = torch.Generator().manual_seed(0)
px = torch.randn((64, 3, 448, 448), generator=g).to(dev)
ppx = torch.randn((64, 3, 448, 448), generator=g).to(dev)
pm = torch.randn((64, 3, 448, 448), generator=g).to(dev)


def run(chunk):
    for s in range(0, 64, chunk):
        model(pixel_values=px[s:s + chunk], prompt_pixel_values=ppx[s:s + chunk], prompt_masks=pm[s:s + chunk],
              embedding_type="instance")


for chunk in (64, 32, 16, 8, 64, 32, 16):
    for _ in range(2):
        run(chunk)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run(chunk)
    e1.record()
    torch.cuda.synchronize()
    print(f"64 tiles as {64 // chunk} x {chunk}: {e0.elapsed_time(e1) / 3:.2f} ms", flush=True)
