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


def run_resid_ln(A, W, bias, h, gamma, beta, eps=1e-6):
    """h (fp32, updated in place) += A W^T + bias; returns the bf16 LayerNorm of the updated rows."""
    M, K = A.shape
    L = _lib.lib()
    ln = torch.empty((M, 1024), dtype=torch.bfloat16, device=A.device)
    scratch = torch.empty(int(L.bseg_gemm_resid_ln_scratch_bytes(M)), dtype=torch.uint8, device=A.device)
    scratch.fill_(0xA5)  # the call zeroes what it needs
    _lib.check(L.bseg_gemm_bf16_resid_ln(_lib.ptr(A), A.stride(0), _lib.ptr(W), M, K, _lib.ptr(bias), _lib.ptr(h),
                                         _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(ln), eps, _lib.ptr(scratch),
                                         _lib.stream_ptr()), "bseg_gemm_bf16_resid_ln")
    torch.cuda.synchronize()
    return ln


@pytest.mark.parametrize("M,K,pairs", [(1568, 1024, 1), (1568, 4096, 1), (1568 * 2 * 8, 1024, 1), (1568 * 7, 4096, 1),
                                       (1568 * 3, 1024, 0), (1568 * 64, 1024, 1)])
def test_gemm_residual_layernorm_epilogue(dev, M, K, pairs):
    """EPI_RESID_LN (HF:modeling_seggpt.py:420-441: residual add, then the LayerNorm that reads the stream next): the
    fp32 stream must equal the plain residual epilogue BIT FOR BIT (same accumulators, same adds), and the fused bf16
    LayerNorm output must match (a) an fp64 LayerNorm of that stream to bf16 rounding and (b) the stand-alone
    layernorm1024 kernel to one bf16 ulp (the statistics are merged in a different order).  Rows carry a large common
    offset (|mean| = 30 std) so that an E[x^2] - mean^2 formulation would fail."""
    g = torch.Generator(device="cpu").manual_seed(M + K)
    A = torch.randn((M, K), generator=g).to(dev).to(torch.bfloat16)
    W = (torch.randn((1024, K), generator=g) / K ** 0.5).to(dev).to(torch.bfloat16)
    bias = torch.randn((1024,), generator=g).to(dev)
    gamma = (1.0 + 0.2 * torch.randn((1024,), generator=g)).to(dev)
    beta = (0.1 * torch.randn((1024,), generator=g)).to(dev)
    h0 = (torch.randn((M, 1024), generator=g) + 30.0 * torch.randn((M, 1), generator=g)).to(dev)
    L = _lib.lib()
    prev = L.bseg_gemm_set_cta_pairs(pairs)
    try:
        h = h0.clone()
        ln = run_resid_ln(A, W, bias, h, gamma, beta)
        h2 = h0.clone()
        ln2 = run_resid_ln(A, W, bias, h2, gamma, beta)
    finally:
        L.bseg_gemm_set_cta_pairs(prev)
    assert torch.equal(h, h2) and torch.equal(ln, ln2)  # deterministic, whichever CTA finishes first
    want_h = run_gemm(A, W, bias) + h0  # (acc + bias) + residual: the epilogue's own order of fp32 additions
    assert torch.equal(h, want_h)
    e, s = report(h, A.float() @ W.float().t() + bias + h0, f"resid_ln stream {M}x{K}")
    assert e <= 2e-3 * s
    sep = torch.empty((M, 1024), dtype=torch.bfloat16, device=dev)
    _lib.check(L.bseg_layernorm1024(_lib.ptr(h), 1024, _lib.ptr(gamma), _lib.ptr(beta), _lib.ptr(sep), 1024, M, 1e-6,
                                    _lib.stream_ptr()), "bseg_layernorm1024")
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(h.double(), (1024,), gamma.double(), beta.double(), 1e-6)
    err = (ln.double() - ref).abs()
    # half a bf16 ulp is 2^-9 relative (allow one); the absolute floor is fp32 rounding of x - mean at |x| ~ 100 std
    tol = ref.abs() * 2.0 ** -8 + 3e-4
    bad = (err > tol).float().mean().item()
    diff_sep = (ln.float() - sep.float()).abs()
    ulp = sep.float().abs() * 2.0 ** -7 + 3e-4
    print(f"[resid_ln {M}x{K} pairs={pairs}] vs fp64 LN: max|err|={err.max().item():.3e} beyond 1 ulp: {bad:.2e}; "
          f"differs from the stand-alone kernel in {(diff_sep > 0).float().mean().item():.2e} of the elements, "
          f"max {(diff_sep / ulp).max().item():.2f} ulp")
    assert bad == 0.0
    assert (diff_sep <= ulp).all()
    assert (diff_sep > 0).float().mean().item() < 2e-2
