"""Runs the fused attention kernel once on a full-size problem (for ncu)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
q = (torch.randn((nseq, 16, 1568, 64), generator=g) * 0.18).to(dev).to(torch.bfloat16)  # pre-scaled by 0.125*log2(e)
k = torch.randn((nseq, 16, 1568, 64), generator=g).to(dev).to(torch.bfloat16)
vt = torch.randn((nseq, 16, 64, 1568), generator=g).to(dev).to(torch.bfloat16)
rel = (torch.randn((176, 64), generator=g) * 0.3 * 8).to(dev).to(torch.bfloat16)  # relcat8
out = torch.empty((nseq, 1568, 1024), dtype=torch.bfloat16, device=dev)
L = _lib.lib()
for _ in range(2):
    _lib.check(L.bseg_attention(_lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(rel), _lib.ptr(out), nseq,
                                _lib.stream_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    _lib.check(L.bseg_attention(_lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(rel), _lib.ptr(out), nseq,
                                _lib.stream_ptr()))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
flops = nseq * 16 * (4.0 * 1568 * 1568 * 64 + 2.0 * 1568 * 84 * 64)
print(f"attention nseq={nseq}: {ms:.3f} ms, {flops / ms / 1e9:.1f} TFLOP/s (algorithmic)")
