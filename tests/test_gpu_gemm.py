"""GPU: tcgen05 GEMM (through the C ABI) against a torch fp32 matmul of the same bf16 operands."""
import ctypes as C

import pytest
import torch

from beach_seg_b200 import _lib

pytestmark = pytest.mark.gpu


def run_gemm(A, W, bias=None, out_bf16=False, gelu=False):
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=A.device)
    _lib.check(_lib.lib().bseg_gemm_bf16(_lib.ptr(A), A.stride(0), _lib.ptr(W), M, N, K, _lib.ptr(bias), _lib.ptr(out),
                                         N, int(out_bf16), int(gelu), _lib.stream_ptr()), "bseg_gemm_bf16")
    torch.cuda.synchronize()
    return out


def report(got, want, tag):
    err = (got.float() - want).abs()
    scale = want.abs().max().item()
    print(f"[{tag}] max|err|={err.max().item():.4e} scale={scale:.3e} mean|err|={err.mean().item():.4e}")
    if err.max().item() > 1e-2 * scale:
        bad = (err > 1e-2 * scale)
        rows = bad.any(dim=1).nonzero().flatten()[:16].tolist()
        cols = bad.any(dim=0).nonzero().flatten()[:16].tolist()
        print(f"[{tag}] bad fraction={bad.float().mean().item():.4f} first bad rows={rows} cols={cols}")
    return err.max().item(), scale


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 256, 64), (128, 128, 256), (300, 256, 128),
                                   (1568, 1024, 1024), (4096 + 17, 3072, 768), (1568 * 3, 1024, 4096)])
def test_gemm_f32_out(dev, M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g).to(dev).to(torch.bfloat16)
    W = torch.randn((N, K), generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(dev)
    got = run_gemm(A, W, bias)
    want = A.float() @ W.float().t() + bias
    e, s = report(got, want, f"gemm {M}x{N}x{K}")
    assert e <= 2e-3 * s


def test_gemm_exact_small_integers(dev):
    """Integer-valued operands: every product and partial sum is exact in fp32, so the result must be bit exact."""
    g = torch.Generator(device="cpu").manual_seed(0)
    A = torch.randint(-4, 5, (640, 512), generator=g).to(dev).to(torch.bfloat16)
    W = torch.randint(-4, 5, (512, 512), generator=g).to(dev).to(torch.bfloat16)
    got = run_gemm(A, W)
    want = A.float() @ W.float().t()
    assert torch.equal(got, want)


def test_gemm_strided_a_and_bf16_gelu(dev):
    g = torch.Generator(device="cpu").manual_seed(5)
    big = torch.randn((777, 4096), generator=g).to(dev).to(torch.bfloat16)
    A = big[:, 1024:2048]  # lda = 4096, like the intermediate-feature slices
    W = (torch.randn((256, 1024), generator=g) * 0.05).to(dev).to(torch.bfloat16)
    bias = torch.randn((256,), generator=g).to(dev)
    got = run_gemm(A, W, bias, out_bf16=True, gelu=True)
    want = torch.nn.functional.gelu(A.float() @ W.float().t() + bias)
    e, s = report(got, want, "gemm gelu bf16")
    assert e <= 1e-2 * s  # bf16 output rounding


def test_gemm_argument_errors(dev):
    A = torch.zeros((128, 100), dtype=torch.bfloat16, device=dev)
    W = torch.zeros((128, 100), dtype=torch.bfloat16, device=dev)
    out = torch.empty((128, 128), dtype=torch.float32, device=dev)
    rc = _lib.lib().bseg_gemm_bf16(_lib.ptr(A), 100, _lib.ptr(W), 128, 128, 100, None, _lib.ptr(out), 128, 0, 0,
                                   _lib.stream_ptr())
    assert rc != 0 and b"multiple of 64" in _lib.lib().bseg_last_error()


@pytest.fixture
def cta_pairs():
    """Runs the body with the CTA-pair GEMM kernel (tcgen05.mma.cta_group::2) selected, then restores the setting."""
    L = _lib.lib()
    prev = L.bseg_gemm_set_cta_pairs(1)
    yield
    L.bseg_gemm_set_cta_pairs(prev)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (128, 256, 64), (300, 256, 128), (1568, 1024, 1024),
                                   (4096 + 17, 3072, 768), (1568 * 3, 1024, 4096), (37 * 256, 4096, 1024)])
def test_gemm_cta_pairs_bit_identical(dev, M, N, K):
    """The CTA-pair kernel accumulates every output in the same order as the one-CTA kernel (same K blocking, fp32
    TMEM accumulators), so the two must agree bit for bit -- including ragged M (TMA zero fill, guarded stores in the
    second CTA of the last pair) -- and both must match the fp32 matmul."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g).to(dev).to(torch.bfloat16)
    W = torch.randn((N, K), generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(dev)
    L = _lib.lib()
    prev = L.bseg_gemm_set_cta_pairs(0)
    try:
        one = run_gemm(A, W, bias)
        one_bf = run_gemm(A, W, bias, out_bf16=True, gelu=True)
        L.bseg_gemm_set_cta_pairs(1)
        two = run_gemm(A, W, bias)
        two_bf = run_gemm(A, W, bias, out_bf16=True, gelu=True)
    finally:
        L.bseg_gemm_set_cta_pairs(prev)
    want = A.float() @ W.float().t() + bias
    e, s = report(two, want, f"gemm pairs {M}x{N}x{K}")
    assert e <= 2e-3 * s
    assert torch.equal(one, two)
    assert torch.equal(one_bf, two_bf)


@pytest.mark.parametrize("M,N,K", [(1568, 1024, 4096), (1568, 1024, 1024), (3136, 3072, 1024), (1568 + 40, 4096, 1024),
                                   (6272, 1024, 4096)])
def test_gemm_small_launch_tiles_bit_identical(dev, M, N, K):
    """bseg_gemm_set_small_tiles: launches of one to four tiles (M = 1568 .. 6272 rows) run on 128 x 128 one-CTA tiles
    when 256-wide tiles would leave most SMs idle.  Same K order, same fp32 accumulators: the outputs must equal the
    256-wide variants bit for bit (fp32 and bf16 + GELU epilogues) and match the fp32 matmul."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K + 1)
    A = torch.randn((M, K), generator=g).to(dev).to(torch.bfloat16)
    W = torch.randn((N, K), generator=g).to(dev).to(torch.bfloat16)
    bias = torch.randn((N,), generator=g).to(dev)
    L = _lib.lib()
    prev = L.bseg_gemm_set_small_tiles(0)
    try:
        wide = run_gemm(A, W, bias)
        wide_bf = run_gemm(A, W, bias, out_bf16=True, gelu=True)
        assert L.bseg_gemm_set_small_tiles(1) == 0
        small = run_gemm(A, W, bias)
        small_bf = run_gemm(A, W, bias, out_bf16=True, gelu=True)
    finally:
        L.bseg_gemm_set_small_tiles(prev)
    want = A.float() @ W.float().t() + bias
    e, s_ = report(small, want, f"gemm small tiles {M}x{N}x{K}")
    assert e <= 2e-3 * s_
    assert torch.equal(wide, small)
    assert torch.equal(wide_bf, small_bf)
