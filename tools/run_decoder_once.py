"""Runs the decoder head kernel (conv3x3 + LN + GELU + 1x1) on the bench shape (64 x 896 x 448 x 64) for timing / ncu."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
H, W = 896, 448
x = torch.randn((B, H, W, 64), generator=g).to(dev).to(torch.bfloat16)
conv_w = (torch.randn((64, 64, 3, 3), generator=g) * 0.05).to(dev)
par = [t.to(dev).contiguous() for t in (torch.randn(64, generator=g) * 0.1, 1 + 0.1 * torch.randn(64, generator=g),
                                        0.1 * torch.randn(64, generator=g), torch.randn((3, 64), generator=g) * 0.2,
                                        torch.randn(3, generator=g) * 0.1)]
L = _lib.lib()
w9 = torch.empty((9, 64, 64), dtype=torch.bfloat16, device=dev)
_lib.check(L.bseg_pack_conv_w9(_lib.ptr(conv_w), _lib.ptr(w9), _lib.stream_ptr()))
pred = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)


def run():
    _lib.check(L.bseg_decoder_head(_lib.ptr(x), _lib.ptr(w9), *[_lib.ptr(p) for p in par], _lib.ptr(pred), B, H, W, 1e-6,
                                   _lib.stream_ptr()))


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flop = B * H * W * (2.0 * 576 * 64 + 2.0 * 64 * 3)
print(f"decoder head B={B}: {ms:.3f} ms, {flop / ms / 1e9:.1f} TFLOP/s")
