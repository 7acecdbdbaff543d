"""Times the layer GEMM shapes of the bench step (64 tiles: M = 100352) with the one-CTA and the CTA-pair kernel.
usage: run_gemm_variants.py [iters]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from beach_seg_b200 import _lib

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
L = _lib.lib()
M = 64 * 1568
shapes = [("proj (resid-like f32 out)", 1024, 1024, False, False), ("qkv-like bf16 out", 3072, 1024, True, False),
          ("lin1 gelu bf16 out", 4096, 1024, True, True), ("lin2 f32 out", 1024, 4096, False, False)]
g = torch.Generator().manual_seed(0)
for name, N, K, bf, gelu in shapes:
    A = (torch.randn((M, K), generator=g) * 0.5).to(dev).to(torch.bfloat16)
    W = (torch.randn((N, K), generator=g) * 0.05).to(dev).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(dev)
    out = torch.empty((M, N), dtype=torch.bfloat16 if bf else torch.float32, device=dev)
    res = {}
    for pairs in (0, 1, 0, 1):
        L.bseg_gemm_set_cta_pairs(pairs)

        def run():
            _lib.check(L.bseg_gemm_bf16(_lib.ptr(A), K, _lib.ptr(W), M, N, K, _lib.ptr(bias), _lib.ptr(out), N, int(bf),
                                        int(gelu), _lib.stream_ptr()))
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res.setdefault(pairs, []).append(ms)
    fl = 2.0 * M * N * K
    one, two = min(res[0]), min(res[1])
    print(f"{name:28s} N={N:5d} K={K:5d}: one CTA {one:.3f} ms {fl / one / 1e9:7.1f} TF | pairs {two:.3f} ms "
          f"{fl / two / 1e9:7.1f} TF | x{one / two:.3f}")
