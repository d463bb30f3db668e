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
        # (one slice per K split, added in a fixed order: no atomics in the forward; dead segment tiles read as zero)
        acc = alloc(ksplit, batch, rows, n_total, dtype=torch.float32, device=A.device)
        call("ptb200_gemm_tn_f16", A, batch, rows, k, lda, rows * lda, taps, shifts, B, n_total, bn, EPI_ATOMIC, None,
             0, None, 0, 0, None, 0, 0, acc, n_total, None, 0, 1, n_total, GEMM_MAX_CTAS[0], ksplit, seg_counts, seg_cap)
        if out is None:
            out = torch.empty(batch, rows, n_total, dtype=torch.float16, device=A.device)
        call("ptb200_bias_act_cast_f16", acc, ksplit, bias, 1 if epi == EPI_BIAS_RELU else 0, batch * rows, n_total, out)
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
# Split-fp16 fp32-equivalent precision (forward and backward): activations / output gradients are [hi | lo | hi]
# triples (3C wide), weights [Wh | Wh | Wl] triples of W * 2^s with alpha = 2^-s (ParamArena.pack_x3). See
# include/ptb200.h. The tensor core accumulates with truncation (measured on B200 with positive operands,
# tests/dev/x3_diag.py: mean signed error -2.4e-4 for one chain over K = 3 x 25088, -2.7e-6 / -1.2e-6 / -4.8e-7 /
# -1.8e-7 with chunks of 8 / 4 / 2 / 1 k-iterations of 64); the kernel promotes its TMEM partial sums to fp32
# registers every X3_CHUNK[0] k-iterations (round 1 did this with split-K atomics through HBM).
X3_CHUNK = [4]
X3_FUSED_WGRAD = [True]  # False: three passes of ptb200_gemm_wgrad_f16 over column slices (tests compare both)
EPI_SPLIT3_MASK, EPI_F32_STORE = 7, 8


def _x3_bn(n_total):
    for bn in (256, 128, 64):
        if n_total % bn == 0:
            return bn
    raise ValueError(f"f16x3 GEMM: n_total={n_total} is not a multiple of 64")


def gemm_tn_x3(A3, B3, alpha, *, taps=1, shifts=None, bn=None, epi=EPI_SPLIT3_RELU, bias=None, w_valid=0, wp=0,
               split=0, n_valid=0, n_total=None, seg=None, ksplit=None, aux=None):
    """A3: [batch, rows, 3K] fp16 triples; B3: [n_rows, taps*3K]. Returns the output triples
    [batch, rows, 3*n_total] (SPLIT3* epilogues; aux = forward activation triples for EPI_SPLIT3_MASK), an fp32
    [batch, rows, n_total] tensor (EPI_F32_STORE) or (d0, d1) fp32 for EPI_F32_SPLIT. ksplit=None: split the
    reduction over CTAs only for skinny problems (fc1 forward)."""
    batch, rows, lda = A3.shape
    k3 = B3.shape[1] // taps
    assert k3 == lda and k3 % 3 == 0
    seg_counts, seg_cap = seg if seg is not None else (None, 0)
    alloc = torch.zeros if seg is not None else torch.empty
    if n_total is None:
        n_total = B3.shape[0]
    dev = A3.device
    if epi == EPI_F32_SPLIT:
        d0 = alloc(batch, rows, split, dtype=torch.float32, device=dev)
        d1 = alloc(batch, rows, n_valid - split, dtype=torch.float32, device=dev)
        call("ptb200_gemm_tn_f16x3", A3, batch, rows, k3, lda, rows * lda, taps, shifts, B3, n_total, bn or n_total,
             epi, bias, 0 if bias is None else bias.numel(), None, 0, 0, w_valid, wp, d0, split, d1, n_valid - split,
             split, n_valid, GEMM_MAX_CTAS[0], 1, seg_counts, seg_cap, float(alpha), None, 0)
        return d0, d1
    if bn is None:
        bn = _x3_bn(n_total)
    nb = 0 if bias is None else bias.numel()
    if ksplit is None:
        ksplit = 1
        if epi in (EPI_SPLIT3_RELU, EPI_SPLIT3):
            tiles = ((rows + 127) // 128) * batch * (n_total // bn)
            k_iters = taps * (k3 // 64)
            if 2 * tiles <= 148 and k_iters >= 128:
                ksplit = max(1, min(148 // tiles, k_iters // 32))
    if ksplit > 1:
        assert epi in (EPI_SPLIT3_RELU, EPI_SPLIT3)
        # one fp32 slice per K split, added in a fixed order by the finishing kernel: no atomics in the forward, so
        # the discrete decisions downstream (top-k, NMS, pseudo-label thresholds) are reproducible run to run.
        # Segment mode skips dead tiles: their rows must read as zero
        acc = (torch.zeros if seg_counts is not None else torch.empty)(ksplit, batch, rows, n_total, dtype=torch.float32,
                                                                       device=dev)
        call("ptb200_gemm_tn_f16x3", A3, batch, rows, k3, lda, rows * lda, taps, shifts, B3, n_total, bn, EPI_F32_STORE,
             None, 0, None, 0, 0, 0, 0, acc, n_total, None, 0, 0, n_total, GEMM_MAX_CTAS[0], ksplit, seg_counts,
             seg_cap, 1.0, None, X3_CHUNK[0])
        out = torch.empty(batch, rows, 3 * n_total, dtype=torch.float16, device=dev)
        call("ptb200_bias_act_split3_f16", acc, ksplit, bias, 1 if epi == EPI_SPLIT3_RELU else 0, float(alpha),
             batch * rows, n_total, wp, w_valid, out)
        return out
    if epi == EPI_F32_STORE:
        d0 = alloc(batch, rows, n_total, dtype=torch.float32, device=dev)
        call("ptb200_gemm_tn_f16x3", A3, batch, rows, k3, lda, rows * lda, taps, shifts, B3, n_total, bn, epi, bias,
             nb, None, 0, 0, w_valid, wp, d0, n_total, None, 0, 0, n_total, GEMM_MAX_CTAS[0], 1, seg_counts, seg_cap,
             float(alpha), None, X3_CHUNK[0])
        return d0
    assert epi in (EPI_SPLIT3_RELU, EPI_SPLIT3, EPI_SPLIT3_MASK)
    assert (aux is not None) == (epi == EPI_SPLIT3_MASK)
    out = alloc(batch, rows, 3 * n_total, dtype=torch.float16, device=dev)
    call("ptb200_gemm_tn_f16x3", A3, batch, rows, k3, lda, rows * lda, taps, shifts, B3, n_total, bn, epi, bias, nb,
         out, 3 * n_total, rows * 3 * n_total, w_valid, wp, None, 0, None, 0, 0, n_total, GEMM_MAX_CTAS[0], 1,
         seg_counts, seg_cap, float(alpha), aux, X3_CHUNK[0])
    return out


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


X3_CONV1_TC = [True]  # False: the fp32 CUDA-core first conv of csrc/split3.cu (tests compare both)


def conv1_u8_x3_tc(images_u8, hw, hmax, wmax, wpack3, bias_table, alpha):
    """Tensor-core f16x3 first conv (raw pixels are exact in fp16; see ParamArena._pack_conv1_x3)."""
    N = hw.shape[0]
    out = torch.empty(N, hmax * (wmax + 1), 192, dtype=torch.float16, device=images_u8.device)
    call("ptb200_conv1_u8_f16x3_tc", images_u8, hw, N, hmax, wmax, images_u8.stride(0), wpack3, bias_table, float(alpha),
         out)
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


def wgrad_x3(G3, X3, out, *, m_total, n_total, taps=1, shifts=None, scale=1.0, bias_out=None, seg=None):
    """f16x3 weight gradient: G3 [batch, rows, 3*m_total] / X3 [batch, rows, 3*n_total] triples;
    out += scale * (Gh'Xh + Gl'Xh + Gh'Xl) as three passes of the MN-major tcgen05 kernel over column slices of the
    triples (the fp32 red.add accumulation into `out` is round-to-nearest); bias_out += scale * colsum(Gh + Gl)."""
    batch, rows, ldg = G3.shape
    ldx = X3.shape[2]
    assert ldg == 3 * m_total and ldx == 3 * n_total
    bn = 256
    while n_total % bn:
        bn //= 2
    tiles = taps * (m_total // 128) * (n_total // bn)
    chunks = ((rows + 63) // 64) * batch
    # one wave of CTAs, and reduction chains of at most 128 chunks (3 x 512 truncating MMAs) per CTA
    ksplit = max(1, min(max(148 // tiles, (chunks + 127) // 128), max(1, chunks // 8)))
    sc, cap = (seg[0], seg[1]) if seg else (None, 0)
    if X3_FUSED_WGRAD[0]:
        call("ptb200_gemm_wgrad_f16x3", G3, ldg, rows * ldg, X3, ldx, rows * ldx, batch, rows, m_total, n_total, taps,
             shifts, out, taps * n_total, float(scale), ksplit, bias_out, sc, cap)
        return out
    # three passes of the plain kernel over column slices of the triples (kept as the cross-check of the fused kernel)
    Gh, Gl = G3[:, :, :m_total], G3[:, :, m_total:2 * m_total]
    Xh, Xl = X3[:, :, :n_total], X3[:, :, n_total:2 * n_total]
    for g, x, b in ((Gh, Xh, bias_out), (Gl, Xh, bias_out), (Gh, Xl, None)):
        call("ptb200_gemm_wgrad_f16", g, ldg, rows * ldg, x, ldx, rows * ldx, batch, rows, m_total, n_total, taps,
             shifts, out, taps * n_total, float(scale), ksplit, b, sc, cap)
    return out


def conv3x3_wgrad_x3(dy: FlatAct, x: FlatAct, out, cout, cin, scale=1.0, bias_out=None):
    return wgrad_x3(dy.t, x.t, out, m_total=cout, n_total=cin, taps=9, shifts=_shifts(x.W + 1), scale=scale,
                    bias_out=bias_out)


def conv3x3_dgrad_x3(dz: FlatAct, wd3, alpha, aux=None):
    """Data gradient of a 3x3 conv in f16x3: aux (forward input triples) selects the fused ReLU backward and a
    triple result; without aux the result is an fp32 [N, rows, Cin] tensor (consumed by the max-pool backward)."""
    Wp = dz.W + 1
    if aux is not None:
        o = gemm_tn_x3(dz.t, wd3, alpha, taps=9, shifts=_shifts(Wp), epi=EPI_SPLIT3_MASK, aux=aux, w_valid=dz.W, wp=Wp)
        return FlatAct(o, dz.H, dz.W)
    return gemm_tn_x3(dz.t, wd3, alpha, taps=9, shifts=_shifts(Wp), epi=EPI_F32_STORE, w_valid=dz.W, wp=Wp)


def maxpool2x2_relu_bwd_x3(x: FlatAct, dpooled_f32):
    N, _, C3 = x.t.shape
    dz = torch.empty_like(x.t)
    call("ptb200_maxpool2x2_relu_bwd_f16x3", x.t, dpooled_f32, dz, N, x.H, x.W, C3 // 3)
    return FlatAct(dz, x.H, x.W)


def pack_grad2_x3(d0, n0, d1, n1, g0, g1, lscale, rows, ld=128):
    out = torch.empty(rows, 3 * ld, dtype=torch.float16, device=d0.device)
    call("ptb200_pack_grad2_f16x3", d0, n0, d1, n1, g0, g1, float(lscale), rows, ld, out)
    return out


def roi_align_bwd_f32(dout_f32, feat_like: FlatAct, rois, counts, cap, scale, pooled):
    N, rows, C3 = feat_like.t.shape
    C = C3 // 3
    dfeat = torch.zeros(N, rows, C, dtype=torch.float32, device=dout_f32.device)
    call("ptb200_roi_align_bwd_f32", dout_f32, N, feat_like.H, feat_like.W, C, rois, counts, cap, float(scale), pooled,
         dfeat)
    return dfeat


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
    wpad = (words + 1) // 2 * 2
    # running `removed` vectors + the mask of one band of candidate rows (the first band holds ~1.25 max_keep rows,
    # the rest of the list is cut into three bands: see ptb200_nms)
    b0 = min(cap, max(1024, max_keep + max_keep // 4))
    b0 = (b0 + 63) // 64
    band = max(b0, max(1, (words - b0 + 2) // 3))
    mask = torch.empty(N * wpad + N * band * 64 * wpad, dtype=torch.int64, device=boxes.device)
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
