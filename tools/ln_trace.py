"""Per-phase cycle counts of the residual+LayerNorm GEMM epilogue (library built with EXTRA=-DBSEG_LN_TRACE):
   python tools/ln_trace.py [M] [K]"""
import sys
import torch
from beach_seg_b200 import _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 100352
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda:0")
L = _lib.lib()
A = torch.randn((M, K), device=dev).to(torch.bfloat16)
W = (torch.randn((1024, K), device=dev) / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(1024, device=dev)
gamma = torch.ones(1024, device=dev)
beta = torch.zeros(1024, device=dev)
h = torch.randn((M, 1024), device=dev)
ln = torch.empty((M, 1024), dtype=torch.bfloat16, device=dev)
scratch = torch.empty(int(L.bseg_gemm_resid_ln_scratch_bytes(M)), dtype=torch.uint8, device=dev)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.bseg_gemm_bf16_resid_ln(_lib.ptr(A), K, _lib.ptr(W), M, K, _lib.ptr(bias), _lib.ptr(h), _lib.ptr(gamma),
                                         _lib.ptr(beta), _lib.ptr(ln), 1e-6, _lib.ptr(scratch), _lib.stream_ptr()), "x")
    e1.record()
    torch.cuda.synchronize()
    print(f"M={M} K={K} run {it}: {e0.elapsed_time(e1):.3f} ms", flush=True)
