"""GPU: decoder head (conv3x3 + LN(C) + GELU + conv1x1 implicit-GEMM kernel) against torch fp32 ops."""
import pytest
import torch
import torch.nn.functional as F

from beach_seg_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W", [(1, 64, 128), (2, 896, 448)])
def test_decoder_head(dev, B, H, W):
    g = torch.Generator().manual_seed(B * 7 + H)
    x = (torch.randn((B, 64, H, W), generator=g)).to(torch.bfloat16).float()
    conv_w = (torch.randn((64, 64, 3, 3), generator=g) * 0.05).to(torch.bfloat16).float()
    conv_b = torch.randn(64, generator=g) * 0.1
    ln_w = 1 + 0.1 * torch.randn(64, generator=g)
    ln_b = 0.1 * torch.randn(64, generator=g)
    head_w = torch.randn((3, 64, 1, 1), generator=g) * 0.2
    head_b = torch.randn(3, generator=g) * 0.1
    y = F.conv2d(x, conv_w, conv_b, padding=1)
    y = F.layer_norm(y.permute(0, 2, 3, 1), (64,), ln_w, ln_b, 1e-6).permute(0, 3, 1, 2)
    want = F.conv2d(F.gelu(y), head_w, head_b)

    L = _lib.lib()
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(dev).to(torch.bfloat16)
    w9 = torch.empty((9, 64, 64), dtype=torch.bfloat16, device=dev)
    _lib.check(L.bseg_pack_conv_w9(_lib.ptr(conv_w.to(dev).contiguous()), _lib.ptr(w9), _lib.stream_ptr()))
    pred = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
    d = lambda t: t.to(dev).contiguous()
    cb, lw, lb, hw, hb = d(conv_b), d(ln_w), d(ln_b), d(head_w.reshape(3, 64)), d(head_b)
    _lib.check(L.bseg_decoder_head(_lib.ptr(x_nhwc), _lib.ptr(w9), _lib.ptr(cb), _lib.ptr(lw), _lib.ptr(lb),
                                   _lib.ptr(hw), _lib.ptr(hb), _lib.ptr(pred), B, H, W, 1e-6, _lib.stream_ptr()),
               "bseg_decoder_head")
    torch.cuda.synchronize()
    err = (pred.cpu() - want).abs()
    print(f"[decoder_head {B}x{H}x{W}] max|err|={err.max().item():.3e} scale={want.abs().max().item():.3e}")
    if err.max().item() > 1e-3:
        bad = (err > 1e-3).any(dim=1)[0]
        ys, xs = bad.nonzero()[:10].t().tolist() if bad.any() else ([], [])
        print("bad frac", bad.float().mean().item(), "first bad (y,x):", list(zip(ys, xs)))
    assert err.max().item() < 2e-3 * max(1.0, want.abs().max().item())
