"""Parity of the tcgen05 weight-gradient kernel against torch fp32 math on fp16-rounded operands."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _wgrad(G, X, taps, shifts, out, scale=1.0, ksplit=0, bias_out=None):
    from probabilisticteacher_b200._lib import lib, ptr, stream_ptr, check
    batch, rows, m = G.shape
    n = X.shape[2]
    sh = (ctypes.c_int * 9)(*(list(shifts) + [0] * (9 - len(shifts))))
    rc = lib().ptb200_gemm_wgrad_f16(
        ptr(G), ctypes.c_int64(m), ctypes.c_int64(rows * m), ptr(X), ctypes.c_int64(n),
        ctypes.c_int64(rows * n), batch, rows, m, n, taps, sh, ptr(out), ctypes.c_int64(taps * n),
        ctypes.c_float(scale), ksplit, ptr(bias_out), None, 0, stream_ptr())
    check(rc, "wgrad")
    torch.cuda.synchronize()


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("rows,m,n,ksplit", [(64, 128, 64, 1), (2048, 1024, 512, 0), (777, 128, 256, 3),
                                             (512, 256, 1024, 0)])
def test_fc_wgrad(cuda, rows, m, n, ksplit):
    g = torch.Generator().manual_seed(rows)
    G = torch.randn(1, rows, m, generator=g).half().to(cuda)
    X = torch.randn(1, rows, n, generator=g).half().to(cuda)
    out = torch.ones(m, n, device=cuda)
    bias = torch.ones(m, device=cuda)
    _wgrad(G, X, 1, [0], out, scale=0.5, ksplit=ksplit, bias_out=bias)
    ref = 1.0 + 0.5 * (G[0].float().t() @ X[0].float())
    assert _rel(out, ref) < 1e-4
    assert _rel(bias, 1.0 + 0.5 * G[0].float().sum(0)) < 1e-4


@pytest.mark.parametrize("H,W,cin,cout", [(12, 21, 64, 128), (50, 83, 512, 512), (7, 130, 128, 256)])
def test_conv_wgrad(cuda, H, W, cin, cout):
    g = torch.Generator().manual_seed(H * W)
    N, Wp = 2, W + 1
    x = torch.randn(N, cin, H, W, generator=g).half()
    gy = torch.randn(N, cout, H, W, generator=g).half()
    xp = torch.zeros(N, H, Wp, cin, dtype=torch.float16)
    xp[:, :, :W] = x.permute(0, 2, 3, 1)
    gp = torch.zeros(N, H, Wp, cout, dtype=torch.float16)
    gp[:, :, :W] = gy.permute(0, 2, 3, 1)
    X = xp.reshape(N, H * Wp, cin).to(cuda)
    G = gp.reshape(N, H * Wp, cout).to(cuda)
    shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    out = torch.zeros(cout, 9 * cin, device=cuda)
    bias = torch.zeros(cout, device=cuda)
    _wgrad(G, X, 9, shifts, out, bias_out=bias)
    xx = x.float().to(cuda).requires_grad_(False)
    w = torch.zeros(cout, cin, 3, 3, device=cuda, requires_grad=True)
    y = torch.nn.functional.conv2d(xx, w, padding=1)
    y.backward(gy.float().to(cuda))
    ref = w.grad.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
    assert _rel(out, ref) < 1e-4
    assert _rel(bias, gy.float().sum((0, 2, 3)).to(cuda)) < 1e-4
