"""Runs the train-augmentation kernels (bseg_train_aug_fwd / _bwd) at the train-step batch (32 x 3 x 448 x 448 fp32, every
optional op forced on) for timing and ncu.  usage: run_augment_once.py [iters]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import synth
from beach_seg_b200.augment import TrainAug
from beach_seg_b200.config import BeachSegConfig

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
B = 32
conf = BeachSegConfig(sharpness_p=1.0, erasing_p=1.0, gauss_p=1.0)
aug = TrainAug(conf, generator=torch.Generator().manual_seed(1))
img = synth.smooth_image(B, 5000).to(dev).requires_grad_(True)
msk = synth.blocky_mask(B, 5001).to(dev)
draw = aug.sample_params(B, 448, 448)
noise = torch.randn(img.shape, device=dev)
d_out = torch.randn(img.shape, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
px = B * 448 * 448


def timed(fn):
    ms = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms)
    return ms[len(ms) // 2], ms[0]


out = {}


def fwd():
    out["o"], _ = aug.apply(img, msk, draw, noise=noise)


def bwd():
    torch.autograd.grad(out["o"], img, d_out, retain_graph=True)


fwd(); bwd()
torch.cuda.synchronize()
for name, fn, bpp in (("fwd", fwd, 62), ("bwd", bwd, 96)):
    med, lo = timed(fn)
    print(f"train_aug {name} B={B}: median {med * 1e3:.1f} us  min {lo * 1e3:.1f} us (host marshalling + 2 launches)  "
          f"{px * bpp / 1e6:.0f} MB -> {px * bpp / med / 1e6:.0f} GB/s")
