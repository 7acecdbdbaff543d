"""Runs the attention backward (prep + dq + dkv kernels) on a full-size problem (for timing and ncu)."""
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib

nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
T = 1568
q = (torch.randn((nseq, 16, T, 64), generator=g) * 0.18).to(dev).to(torch.bfloat16)  # pre-scaled by 0.125*log2(e)
k = torch.randn((nseq, 16, T, 64), generator=g).to(dev).to(torch.bfloat16)
vt = torch.randn((nseq, 16, 64, T), generator=g).to(dev).to(torch.bfloat16)
rel = (torch.randn((176, 64), generator=g) * 0.3 * 8).to(dev).to(torch.bfloat16)  # relcat8
do = torch.randn((nseq, T, 1024), generator=g).to(dev).to(torch.bfloat16)
out = torch.empty((nseq, T, 1024), dtype=torch.bfloat16, device=dev)
lse = torch.empty((nseq, 16, T), dtype=torch.float32, device=dev)
dqkv = torch.empty((nseq * T, 3072), dtype=torch.bfloat16, device=dev)
L = _lib.lib()
_lib.check(L.bseg_attention_fwd_lse(_lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(rel), _lib.ptr(out), _lib.ptr(lse),
                                    nseq, _lib.stream_ptr()))
nbytes = int(L.bseg_attention_bwd_scratch_bytes(nseq))
scratch = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
base = (scratch.data_ptr() + 255) // 256 * 256


def run():
    _lib.check(L.bseg_attention_bwd(_lib.ptr(q), _lib.ptr(k), _lib.ptr(vt), _lib.ptr(out), _lib.ptr(do), _lib.ptr(lse),
                                    _lib.ptr(rel), _lib.ptr(dqkv), nseq, C.c_void_p(base), C.c_size_t(nbytes),
                                    _lib.stream_ptr()))


for _ in range(2):
    run()
torch.cuda.synchronize()
L.bseg_profile_enable(1)
for _ in range(iters):
    run()
torch.cuda.synchronize()
n = 9
pms, pl, pw, pb = (C.c_double * n)(), (C.c_longlong * n)(), (C.c_double * n)(), (C.c_double * n)()
L.bseg_profile_collect(pms, pl, pw, pb)
print(f"attention bwd nseq={nseq}: attention-category {pms[1] / iters:.3f} ms/iter ({pl[1] // iters} launches), "
      f"prep (elementwise) {pms[7] / iters:.3f} ms/iter; {pw[1] / (pms[1] * 1e-3) / 1e12:.1f} TFLOP/s")
