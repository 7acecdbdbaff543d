"""Bitwise checks of the forward attention kernel: run-to-run determinism and independence of a sequence's result from
the batch it is launched in."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
nseq = 64
q = (torch.randn((nseq, 16, 1568, 64), generator=g) * 1.5 * 0.18).to(dev).to(torch.bfloat16)  # pre-scaled
k = (torch.randn((nseq, 16, 1568, 64), generator=g) * 1.5).to(dev).to(torch.bfloat16)
vt = torch.randn((nseq, 16, 64, 1568), generator=g).to(dev).to(torch.bfloat16)
rel = (torch.randn((176, 64), generator=g) * 0.3 * 8).to(dev).to(torch.bfloat16)  # relcat8
L = _lib.lib()


def run(qq, kk, vv):
    n = qq.shape[0]
    out = torch.empty((n, 1568, 1024), dtype=torch.bfloat16, device=dev)
    _lib.check(L.bseg_attention(_lib.ptr(qq), _lib.ptr(kk), _lib.ptr(vv), _lib.ptr(rel), _lib.ptr(out), n,
                                _lib.stream_ptr()))
    torch.cuda.synchronize()
    return out


ref = run(q, k, vt)
bad = 0
for it in range(10):
    o = run(q, k, vt)
    d = (o.view(torch.int16) != ref.view(torch.int16))
    if d.any():
        bad += 1
        idx = d.nonzero()[:5].tolist()
        print(f"run {it}: {int(d.sum())} elements differ, first {idx}, max abs diff "
              f"{(o.float() - ref.float()).abs().max().item():.3e}")
print("run-to-run mismatching runs:", bad)
for lo, hi in ((0, 21), (21, 42), (42, 64), (5, 6)):
    o = run(q[lo:hi].contiguous(), k[lo:hi].contiguous(), vt[lo:hi].contiguous())
    d = (o.view(torch.int16) != ref[lo:hi].view(torch.int16))
    print(f"subset [{lo},{hi}): {int(d.sum())} elements differ"
          + (f", first {d.nonzero()[:5].tolist()}" if d.any() else ""))
