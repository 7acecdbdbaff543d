"""lin1-shaped GEMM (M = 64 x 1568, N = 4096, K = 1024) with the plain bf16 epilogue and with the GELU epilogue, and the
lin2 / qkv shapes for reference: which part of the lin1 time is the activation?"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib

dev = torch.device("cuda:0")
L = _lib.lib()
g = torch.Generator().manual_seed(0)


def run(M, N, K, out_bf16, gelu, reps=20):
    A = (torch.randn((M, K), generator=g) * 0.5).to(dev).to(torch.bfloat16)
    W = (torch.randn((N, K), generator=g) * 0.05).to(dev).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(dev)
    out = torch.empty((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)

    def call():
        _lib.check(L.bseg_gemm_bf16(_lib.ptr(A), A.stride(0), _lib.ptr(W), M, N, K, _lib.ptr(bias), _lib.ptr(out), N,
                                    int(out_bf16), int(gelu), _lib.stream_ptr()))

    for _ in range(5):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"M={M} N={N} K={K} out={'bf16' if out_bf16 else 'f32 '} gelu={int(gelu)}: {ms:.3f} ms  "
          f"{2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)


M = 64 * 1568
for _ in range(2):
    run(M, 4096, 1024, True, False)
    run(M, 4096, 1024, True, True)
    run(M, 1024, 4096, False, False)
    run(M, 3072, 1024, True, False)
