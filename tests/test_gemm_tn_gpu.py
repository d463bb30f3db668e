"""Parity of the tcgen05 implicit-GEMM kernel (through the C ABI) against torch fp32 math on the
same fp16-rounded operands. Tolerance: fp32 accumulation, fp16 output rounding -> 2e-3 rel."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, B, bias, epi, bn, taps=1, shifts=None, wv=0, wp=0, aux=None, n_valid=0, split=0,
          max_ctas=0):
    from probabilisticteacher_b200._lib import lib, ptr, stream_ptr, check
    batch, rows, lda = A.shape
    n_total = B.shape[0]
    k = B.shape[1] // taps
    sh = (ctypes.c_int * 9)(*(list(shifts) + [0] * (9 - len(shifts)))) if shifts else None
    if epi == 2:
        d0 = torch.full((batch, rows, split), float("nan"), device=A.device)
        d1 = torch.full((batch, rows, n_valid - split), float("nan"), device=A.device)
        D = None
    else:
        D = torch.full((batch, rows, n_total), float("nan"), device=A.device, dtype=torch.float16)
        d0 = d1 = None
    rc = lib().ptb200_gemm_tn_f16(
        ptr(A), batch, rows, k, ctypes.c_int64(lda), ctypes.c_int64(rows * lda), taps, sh, ptr(B),
        n_total, bn, epi, ptr(bias), 0 if bias is None else bias.numel(), ptr(D),
        ctypes.c_int64(n_total), ctypes.c_int64(rows * n_total), ptr(aux), wv, wp, ptr(d0),
        split, ptr(d1), n_valid - split, split, n_valid, max_ctas, 1, None, 0, stream_ptr())
    check(rc, "gemm_tn")
    torch.cuda.synchronize()
    return (d0, d1) if epi == 2 else D


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("rows,k,n,bn", [(128, 64, 64, 64), (300, 256, 128, 128), (1000, 512, 512, 256),
                                         (77, 1024, 1024, 256), (4097, 128, 192, 64)])
def test_plain_gemm_bias_relu(cuda, rows, k, n, bn):
    g = torch.Generator().manual_seed(rows + k)
    A = torch.randn(1, rows, k, generator=g).half().to(cuda)
    B = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(cuda)
    bias = torch.randn(n, generator=g).to(cuda)
    D = _gemm(A, B, bias, 0, bn)
    ref = torch.relu(A[0].float() @ B.float().t() + bias)
    assert _rel(D[0], ref) < 2e-3
    D = _gemm(A, B, bias, 1, bn)
    ref = A[0].float() @ B.float().t() + bias
    assert _rel(D[0], ref) < 2e-3


def test_persistent_multi_tile_few_ctas(cuda):
    g = torch.Generator().manual_seed(5)
    A = torch.randn(2, 1500, 128, generator=g).half().to(cuda)
    B = (torch.randn(256, 128, generator=g) / 11).half().to(cuda)
    bias = torch.randn(256, generator=g).to(cuda)
    D = _gemm(A, B, bias, 1, 64, max_ctas=3)
    ref = A.float() @ B.float().t() + bias
    assert _rel(D, ref) < 2e-3


def test_f32_split(cuda):
    g = torch.Generator().manual_seed(7)
    A = torch.randn(2, 333, 512, generator=g).half().to(cuda)
    B = torch.zeros(96, 512)
    B[:81] = torch.randn(81, 512, generator=g) / 22
    B = B.half().to(cuda)
    bias = torch.randn(81, generator=g).to(cuda)
    d0, d1 = _gemm(A, B, bias, 2, 96, n_valid=81, split=9)
    ref = A.float() @ B.float().t()[:, :81] + bias
    assert _rel(d0, ref[..., :9]) < 1e-4
    assert _rel(d1, ref[..., 9:]) < 1e-4


@pytest.mark.parametrize("H,W,cin,cout,bn", [(20, 37, 64, 64, 64), (50, 83, 512, 512, 256),
                                             (9, 300, 128, 256, 128)])
def test_conv3x3_flat_padded(cuda, H, W, cin, cout, bn):
    """3x3 conv, stride 1, pad 1, as 9 row-shifted taps over [N][H*(W+1)][C] activations."""
    g = torch.Generator().manual_seed(H * W)
    N, Wp = 2, W + 1
    x = torch.randn(N, cin, H, W, generator=g).half()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).half()
    bias = torch.randn(cout, generator=g)
    xp = torch.zeros(N, H, Wp, cin, dtype=torch.float16)
    xp[:, :, :W] = x.permute(0, 2, 3, 1)
    A = xp.reshape(N, H * Wp, cin).to(cuda)
    Bw = w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous().to(cuda)
    shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    D = _gemm(A, Bw, bias.to(cuda), 0, bn, taps=9, shifts=shifts, wv=W, wp=Wp)
    ref = torch.relu(torch.nn.functional.conv2d(x.float().to(cuda), w.float().to(cuda),
                                                bias.to(cuda), padding=1))
    got = D.reshape(N, H, Wp, cout)
    assert got[:, :, W:].abs().max().item() == 0.0  # pad column stays zero
    assert _rel(got[:, :, :W].permute(0, 3, 1, 2), ref) < 2e-3


def test_mask_epilogue(cuda):
    g = torch.Generator().manual_seed(11)
    A = torch.randn(1, 700, 128, generator=g).half().to(cuda)
    B = (torch.randn(128, 128, generator=g) / 11).half().to(cuda)
    aux = torch.randn(1, 700, 128, generator=g).half().to(cuda)
    D = _gemm(A, B, None, 3, 128, aux=aux)
    ref = (A.float() @ B.float().t()) * (aux > 0)
    assert _rel(D, ref) < 2e-3


def test_split_k_fc(cuda):
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(21)
    A = torch.randn(1, 300, 2048, generator=g).half().to(cuda)
    B = (torch.randn(512, 2048, generator=g) / 45).half().to(cuda)
    bias = torch.randn(512, generator=g).to(cuda)
    D = ops.gemm_tn(A, B, epi=ops.EPI_BIAS_RELU, bias=bias, ksplit=4)
    torch.cuda.synchronize()
    ref = torch.relu(A[0].float() @ B.float().t() + bias)
    assert _rel(D[0], ref) < 2e-3


def test_segment_skipping(cuda):
    """Fixed-capacity roi buffers: tiles / chunks without a live row are skipped; live rows are exact."""
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(31)
    cap, nseg, K, N = 600, 3, 256, 256
    counts = torch.tensor([130, 0, 600], dtype=torch.int32, device=cuda)
    A = torch.randn(1, nseg * cap, K, generator=g).half().to(cuda)
    B = (torch.randn(N, K, generator=g) / 16).half().to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    D = ops.gemm_tn(A, B, epi=ops.EPI_BIAS_RELU, bias=bias, seg=(counts, cap))
    ref = torch.relu(A[0].float() @ B.float().t() + bias)
    live = torch.zeros(nseg * cap, dtype=torch.bool, device=cuda)
    live[:130] = True
    live[2 * cap:] = True
    assert _rel(D[0][live], ref[live]) < 2e-3
    assert D[0][cap + 128:2 * cap - 128].abs().max() == 0  # tiles entirely inside the empty segment: untouched
    # weight gradient with the same segment structure
    G = torch.randn(1, nseg * cap, 128, generator=g).half().to(cuda)
    G[0][~live] = 0
    out = torch.zeros(128, K, device=cuda)
    bsum = torch.zeros(128, device=cuda)
    ops.wgrad(G, A, out, bias_out=bsum, seg=(counts, cap))
    torch.cuda.synchronize()
    assert _rel(out, G[0].float().t() @ A[0].float()) < 1e-4
    assert _rel(bsum, G[0].float().sum(0)) < 1e-4
