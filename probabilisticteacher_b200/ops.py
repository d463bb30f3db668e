"""Torch-tensor wrappers over the C ABI (include/ptb200.h). Tensors only carry device memory; all
arithmetic happens inside libptb200.so. No CPU fallback: every function needs CUDA tensors."""
from collections import namedtuple

import torch

from ._lib import call

EPI_BIAS_RELU, EPI_BIAS, EPI_F32_SPLIT, EPI_MASK, EPI_ATOMIC = 0, 1, 2, 3, 4
EPI_SPLIT3_RELU, EPI_SPLIT3 = 5, 6  # f16x3 parity precision only

# fp16 activation in the flattened right-padded layout: t is [N, H*(W+1), C]
FlatAct = namedtuple("FlatAct", ["t", "H", "W"])

I32 = torch.int32
# upper bound on the CTAs of the persistent GEMM kernel (0 = one per SM). The concurrent step leaves a
# few SMs free so that latency-bound kernels of another stream never wait behind persistent GEMM CTAs.
GEMM_MAX_CTAS = [0]
U32 = torch.int32  # uint32 payloads are carried in int32 tensors (bit patterns only)


def _shifts(Wp):
    return [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]


def flat_zeros(N, H, W, C, device):
    return FlatAct(torch.zeros(N, H * (W + 1), C, dtype=torch.float16, device=device), H, W)


def to_flat(x_nchw):
    """NCHW float tensor -> FlatAct (test helper)."""
    N, C, H, W = x_nchw.shape
    t = torch.zeros(N, H, W + 1, C, dtype=torch.float16, device=x_nchw.device)
    t[:, :, :W] = x_nchw.permute(0, 2, 3, 1).to(torch.float16)
    return FlatAct(t.reshape(N, H * (W + 1), C), H, W)


def from_flat(a):
    """FlatAct -> NCHW fp16 tensor (test helper)."""
    N, _, C = a.t.shape
    return a.t.reshape(N, a.H, a.W + 1, C)[:, :, :a.W].permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------------ GEMMs
def gemm_tn(A, B, *, taps=1, shifts=None, bn=None, epi=EPI_BIAS, bias=None, out=None, aux=None,
            w_valid=0, wp=0, d0=None, d1=None, split=0, n_valid=0, n_total=None, ksplit=1, seg=None):
    """A: [batch, rows, K] fp16; B: [n_rows, taps*K] fp16. Returns out (fp16 [batch, rows, n_total]) or
    (d0, d1) for the fp32 split epilogue."""
    batch, rows, lda = A.shape
    k = B.shape[1] // taps
    seg_counts, seg_cap = seg if seg is not None else (None, 0)
    alloc = torch.zeros if seg is not None else torch.empty  # skipped tiles are never written
    if n_total is None:
        n_total = B.shape[0]
    if bn is None:
        bn = 256
        while n_total % bn:
            bn //= 2
    if ksplit > 1:
        # skinny problem: split the reduction over CTAs, reduce partials in fp32, then bias/act/cast
        assert epi in (EPI_BIAS_RELU, EPI_BIAS) and aux is None
        acc = torch.zeros(batch, rows, n_total, dtype=torch.float32, device=A.device)
        call("ptb200_gemm_tn_f16", A, batch, rows, k, lda, rows * lda, taps, shifts, B, n_total, bn, EPI_ATOMIC, None,
             0, None, 0, 0, None, 0, 0, acc, n_total, None, 0, 0, n_total, GEMM_MAX_CTAS[0], ksplit, seg_counts, seg_cap)
        if out is None:
            out = torch.empty(batch, rows, n_total, dtype=torch.float16, device=A.device)
        call("ptb200_bias_act_cast_f16", acc, bias, 1 if epi == EPI_BIAS_RELU else 0, batch * rows, n_total, out)
        return out
    if epi == EPI_F32_SPLIT:
        if d0 is None:
            d0 = alloc(batch, rows, split, dtype=torch.float32, device=A.device)
        if d1 is None:
            d1 = alloc(batch, rows, n_valid - split, dtype=torch.float32, device=A.device)
        ld_d, dbs = 0, 0
    else:
        if out is None:
            out = alloc(batch, rows, n_total, dtype=torch.float16, device=A.device)
        ld_d, dbs = out.shape[2], out.shape[1] * out.shape[2]
    call("ptb200_gemm_tn_f16", A, batch, rows, k, lda, rows * lda, taps, shifts, B, n_total, bn, epi, bias,
         0 if bias is None else bias.numel(), out, ld_d, dbs, aux, w_valid, wp, d0, split, d1,
         n_valid - split, split, n_valid, GEMM_MAX_CTAS[0], 1, seg_counts, seg_cap)
    return (d0, d1) if epi == EPI_F32_SPLIT else out


def conv3x3(x: FlatAct, w_packed, bias, relu=True, aux=None, out=None):
    """3x3 / stride 1 / pad 1 convolution over a FlatAct. w_packed: fp16 [Cout, 9*Cin] ([co][ky][kx][ci]).
    aux != None selects the ReLU-mask epilogue (data-gradient path)."""
    Wp = x.W + 1
    epi = EPI_MASK if aux is not None else (EPI_BIAS_RELU if relu else EPI_BIAS)
    o = gemm_tn(x.t, w_packed, taps=9, shifts=_shifts(Wp), epi=epi, bias=bias, aux=aux, w_valid=x.W, wp=Wp,
                out=out)
    return FlatAct(o, x.H, x.W)


# ------------------------------------------------------------------------------------------ f16x3
# Split-fp16 parity precision (forward only): activations are [hi | lo | hi] triples (3C wide), weights
# [Wh | Wh | Wl] triples of W * 2^s with alpha = 2^-s (ParamArena.pack_x3). See include/ptb200.h.
# longest chain of tensor-core accumulations (k-iterations of 64 = 4 MMAs each) before the partial sum is
# promoted to a round-to-nearest fp32 add: the MMA accumulates with truncation. Measured on B200 with positive
# operands (tests/dev/x3_diag.py): mean signed error -2.4e-4 for one chain over K = 3 x 25088, -2.7e-6 / -1.2e-6 /
# -4.8e-7 / -1.8e-7 with chunks of 8 / 4 / 2 / 1 k-iterations; the bias compounds through the 16 stacked layers.
X3_MAX_K_ITERS = [2]


def gemm_tn_x3(A3, B3, alpha, *, taps=1, shifts=None, bn=None, epi=EPI_SPLIT3_RELU, bias=None, w_valid=0, wp=0,
               split=0, n_valid=0, n_total=None, seg=None, ksplit=None):
    """A3: [batch, rows, 3K] fp16 triples; B3: [n_rows, taps*3K]. Returns the output triples
    [batch, rows, 3*n_total] or (d0, d1) fp32 for EPI_F32_SPLIT. ksplit=None chooses the K chunking from
    X3_MAX_K_ITERS (triple epilogues only)."""
    batch, rows, lda = A3.shape
    k3 = B3.shape[1] // taps
    assert k3 == lda and k3 % 3 == 0
    seg_counts, seg_cap = seg if seg is not None else (None, 0)
    alloc = torch.zeros if seg is not None else torch.empty
    if n_total is None:
        n_total = B3.shape[0]
    if bn is None:
        bn = 256
        while n_total % bn:
            bn //= 2
    out = d0 = d1 = None
    ld_d = dbs = 0
    if ksplit is None:
        k_iters = taps * (k3 // 64)
        ksplit = (k_iters + X3_MAX_K_ITERS[0] - 1) // X3_MAX_K_ITERS[0] if epi in (EPI_SPLIT3_RELU, EPI_SPLIT3) else 1
    if ksplit > 1:
        assert epi in (EPI_SPLIT3_RELU, EPI_SPLIT3)
        acc = torch.zeros(batch, rows, n_total, dtype=torch.float32, device=A3.device)
        call("ptb200_gemm_tn_f16x3", A3, batch, rows, k3, lda, rows * lda, taps, shifts, B3, n_total, bn, EPI_ATOMIC,
             None, 0, None, 0, 0, 0, 0, acc, n_total, None, 0, 0, n_total, GEMM_MAX_CTAS[0], ksplit, seg_counts,
             seg_cap, 1.0)
        out = torch.empty(batch, rows, 3 * n_total, dtype=torch.float16, device=A3.device)
        call("ptb200_bias_act_split3_f16", acc, bias, 1 if epi == EPI_SPLIT3_RELU else 0, float(alpha), batch * rows,
             n_total, wp, w_valid, out)
        return out
    if epi == EPI_F32_SPLIT:
        d0 = alloc(batch, rows, split, dtype=torch.float32, device=A3.device)
        d1 = alloc(batch, rows, n_valid - split, dtype=torch.float32, device=A3.device)
    else:
        out = alloc(batch, rows, 3 * n_total, dtype=torch.float16, device=A3.device)
        ld_d, dbs = 3 * n_total, rows * 3 * n_total
    call("ptb200_gemm_tn_f16x3", A3, batch, rows, k3, lda, rows * lda, taps, shifts, B3, n_total, bn, epi, bias,
         0 if bias is None else bias.numel(), out, ld_d, dbs, w_valid, wp, d0, split, d1, n_valid - split, split,
         n_valid, GEMM_MAX_CTAS[0], 1, seg_counts, seg_cap, float(alpha))
    return (d0, d1) if epi == EPI_F32_SPLIT else out


def conv3x3_x3(x: FlatAct, w3, alpha, bias):
    """3x3 conv + bias + ReLU over f16x3 triples. w3: [Cout, 9 * 3Cin]."""
    Wp = x.W + 1
    o = gemm_tn_x3(x.t, w3, alpha, taps=9, shifts=_shifts(Wp), epi=EPI_SPLIT3_RELU, bias=bias, w_valid=x.W, wp=Wp)
    return FlatAct(o, x.H, x.W)


def split3_pack(src_f32, k, scale=1.0, order=0):
    """fp32 [..., k] -> fp16 [rows, 3k] triples (order 0: activation [hi|lo|hi]; 1: weight [hi|hi|lo])."""
    rows = src_f32.numel() // k
    dst = torch.empty(rows, 3 * k, dtype=torch.float16, device=src_f32.device)
    call("ptb200_split3_pack_f16", src_f32.contiguous(), dst, rows, k, float(scale), order)
    return dst


def split3_unpack(t3, k):
    """fp16 triples [..., 3k] -> fp32 [rows, k] (hi + lo)."""
    rows = t3.numel() // (3 * k)
    dst = torch.empty(rows, k, dtype=torch.float32, device=t3.device)
    call("ptb200_split3_unpack_f32", t3.contiguous(), dst, rows, k)
    return dst


def conv1_u8_x3(images_u8, hw, hmax, wmax, mean, std, w_f32, bias):
    N = hw.shape[0]
    out = torch.empty(N, hmax * (wmax + 1), 192, dtype=torch.float16, device=images_u8.device)
    call("ptb200_conv1_u8_f16x3", images_u8, hw, N, hmax, wmax, images_u8.stride(0), list(mean), list(std), w_f32,
         bias, out)
    return FlatAct(out, hmax, wmax)


def maxpool2x2_x3(x: FlatAct):
    N, _, C3 = x.t.shape
    Ho, Wo = x.H // 2, x.W // 2
    out = torch.empty(N, Ho * (Wo + 1), C3, dtype=torch.float16, device=x.t.device)
    call("ptb200_maxpool2x2_f16x3", x.t, out, N, x.H, x.W, C3 // 3)
    return FlatAct(out, Ho, Wo)


def roi_align_fwd_x3(feat: FlatAct, rois, counts, cap, scale, pooled):
    N, _, C3 = feat.t.shape
    out = torch.empty(N * cap, pooled * pooled * C3, dtype=torch.float16, device=feat.t.device)
    call("ptb200_roi_align_fwd_f16x3", feat.t, N, feat.H, feat.W, C3 // 3, rois, counts, cap, float(scale), pooled, out)
    return out


def wgrad(G, X, out, *, taps=1, shifts=None, scale=1.0, ksplit=0, m_total=None, n_total=None, bias_out=None,
          seg=None):
    """out[m][t*n + n'] += scale * sum G[b][p][m] X[b][p+shift_t][n'];  bias_out[m] += scale * sum G[b][p][m]."""
    batch, rows, ldg = G.shape
    ldx = X.shape[2]
    m_total = m_total or ldg
    n_total = n_total or ldx
    call("ptb200_gemm_wgrad_f16", G, ldg, rows * ldg, X, ldx, rows * ldx, batch, rows, m_total, n_total, taps,
         shifts, out, taps * n_total, float(scale), ksplit, bias_out, seg[0] if seg else None, seg[1] if seg else 0)
    return out


def conv3x3_wgrad(dy: FlatAct, x: FlatAct, out, scale=1.0, bias_out=None):
    return wgrad(dy.t, x.t, out, taps=9, shifts=_shifts(x.W + 1), scale=scale, bias_out=bias_out)


# ------------------------------------------------------------------------------------------ elementwise
def preprocess_im2col(images_u8, hw, hmax, wmax, mean, std):
    """images_u8: uint8 [N, 3, hmax, wmax] (device) when all images share a size, else a flat buffer with
    image_stride; hw: int32 [N, 2] device. Returns FlatAct with C = 64."""
    N = hw.shape[0]
    out = torch.empty(N, hmax * (wmax + 1), 64, dtype=torch.float16, device=images_u8.device)
    call("ptb200_preprocess_im2col", images_u8, hw, N, hmax, wmax, images_u8.stride(0), list(mean), list(std), out)
    return FlatAct(out, hmax, wmax)


def conv1_u8(images_u8, hw, hmax, wmax, mean, std, wpack, bias):
    """Pre-processing + first VGG conv fused: uint8 images -> FlatAct with C = 64 (bias + ReLU applied)."""
    N = hw.shape[0]
    out = torch.empty(N, hmax * (wmax + 1), 64, dtype=torch.float16, device=images_u8.device)
    call("ptb200_conv1_u8_f16", images_u8, hw, N, hmax, wmax, images_u8.stride(0), list(mean), list(std), wpack,
         bias, out)
    return FlatAct(out, hmax, wmax)


def maxpool2x2(x: FlatAct):
    N, _, C = x.t.shape
    Ho, Wo = x.H // 2, x.W // 2
    out = torch.empty(N, Ho * (Wo + 1), C, dtype=torch.float16, device=x.t.device)
    call("ptb200_maxpool2x2_f16", x.t, out, N, x.H, x.W, C)
    return FlatAct(out, Ho, Wo)


def maxpool2x2_relu_bwd(x: FlatAct, dpooled: FlatAct):
    N, _, C = x.t.shape
    dz = torch.empty_like(x.t)
    call("ptb200_maxpool2x2_relu_bwd_f16", x.t, dpooled.t, dz, N, x.H, x.W, C)
    return FlatAct(dz, x.H, x.W)


def colsum(x2d, out, scale=1.0, c=None):
    rows, ld = x2d.shape
    call("ptb200_colsum_f16", x2d, rows, c or ld, ld, float(scale), out)


def segmented_sort(keys, vals, seg_len=None, max_len=None, begin_bit=0, end_bit=32):
    """In-place stable ascending sort of each row of keys/vals ([segments, stride] int32 bit patterns)."""
    segs, stride = keys.shape
    kt = torch.empty_like(keys)
    vt = torch.empty_like(vals)
    call("ptb200_segmented_sort_u32", keys, vals, kt, vt, segs, stride, seg_len, max_len or stride, begin_bit,
         end_bit)


def nms(boxes, order, counts, thresh, max_keep, class_mod=0):
    """boxes: fp32 [N, nbox, 4]; order: int32 [N, cap] candidate indices in descending-score order;
    counts: int32 [N]. Returns keep_idx [N, max_keep] (positions in `order`), keep_count [N]."""
    N, cap = order.shape
    words = (cap + 63) // 64
    mask = torch.empty(N * cap * ((words + 1) // 2 * 2), dtype=torch.int64, device=boxes.device)
    keep_idx = torch.zeros(N, max_keep, dtype=I32, device=boxes.device)
    keep_count = torch.zeros(N, dtype=I32, device=boxes.device)
    call("ptb200_nms", boxes, boxes.shape[1], order, cap, counts, N, cap, float(thresh), class_mod, max_keep, mask,
         keep_idx, keep_count)
    return keep_idx, keep_count


def roi_align_fwd(feat: FlatAct, rois, counts, cap, scale, pooled):
    N, _, C = feat.t.shape
    out = torch.empty(N * cap, pooled * pooled * C, dtype=torch.float16, device=feat.t.device)
    call("ptb200_roi_align_fwd_f16", feat.t, N, feat.H, feat.W, C, rois, counts, cap, float(scale), pooled, out)
    return out


def roi_align_bwd(dout, feat_like: FlatAct, rois, counts, cap, scale, pooled):
    N, rows, C = feat_like.t.shape
    dfeat = torch.zeros(N, rows, C, dtype=torch.float32, device=dout.device)
    call("ptb200_roi_align_bwd_f16", dout, N, feat_like.H, feat_like.W, C, rois, counts, cap, float(scale), pooled,
         dfeat)
    return dfeat
