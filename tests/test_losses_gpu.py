"""GPU: the four fused loss kernels (csrc/losses.cu) IN ISOLATION -- identical fp32 logits / deltas / labels / targets
go to `ptb200_{rpn,roi}_loss_{sup,unsup}` and to the oracle's loss functions (oracle/pt_oracle.py, pinned to the
reference's own functions by tests/test_oracle_golden.py); the oracle runs in fp64 with autograd as the referee.

Compared: the two loss values of each kernel and their analytic unit gradients (`dlogits`, `ddeltas`, `dscores`,
`danchor_wh`) -- tolerance 1e-5 of each tensor's max magnitude (the kernels are fp32: expf / logf / powf).
Reference: pt/modeling/proposal_generator/rpn.py:191-361, pt/modeling/roi_heads/fast_rcnn.py:179-336,
pt/modeling/box_regression.py:33-35,66-99,142-201. Variants: EFL on/off, tau 0.5 / 0.25, lambda off 1, danchor on/off,
an image without positives, the empty (NaN) edge of the unsupervised ROI losses."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _anchors(O, H, W, wh):
    return O.grid_anchors(O.differentiable_cell_anchors(wh), H, W, 16, 0.5)


def _flat_from_nr(x_nr, N, H, W, A, width, fill):
    """[N, R(, k)] in anchor order r = (y*W + x)*A + a -> flat rows [N*H*(W+1), A*width]; pad column rows = fill."""
    x = x_nr.reshape(N, H, W, A * width)
    out = torch.full((N, H, W + 1, A * width), fill, dtype=x.dtype)
    out[:, :, :W] = x
    return out.reshape(N * H * (W + 1), A * width).contiguous()


def _nr_from_flat(x_flat, N, H, W, A, width):
    return x_flat.reshape(N, H, W + 1, A * width)[:, :, :W].reshape(N, H * W * A, width)


# ------------------------------------------------------------------------------------------ RPN supervised
@pytest.mark.parametrize("seed,empty_image", [(0, False), (1, True)])
def test_rpn_loss_sup_kernel(cuda, seed, empty_image):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200._lib import call
    g = torch.Generator().manual_seed(seed)
    N, H, W, A, cap = 2, 6, 9, 9, 16
    R = H * W * A
    wh = torch.tensor([[181., 90.], [128., 128.], [90., 181.], [362., 181.], [256., 256.], [181., 362.], [724., 362.],
                       [512., 512.], [362., 724.]])
    anchors = _anchors(O, H, W, wh)
    logits = torch.randn(N, R, generator=g) * 2
    deltas = torch.randn(N, R, 8, generator=g)
    labels = torch.randint(-1, 2, (N, R), generator=g)
    if empty_image:
        labels[1][labels[1] == 1] = 0  # an image whose sampled anchors are all negatives
    matched = torch.randint(0, cap, (N, R), generator=g)
    xy = torch.rand(N, cap, 2, generator=g) * torch.tensor([W * 16 * 0.7, H * 16 * 0.7])
    gt = torch.cat([xy, xy + 8 + torch.rand(N, cap, 2, generator=g) * 60], -1)
    bpi = 256
    norm = 1.0 / (bpi * N)
    # ---- referee
    lo = logits.double().requires_grad_(True)
    do = deltas.double().requires_grad_(True)
    ref = O.rpn_losses(anchors.double(), lo, [labels[n] for n in range(N)], do,
                       [gt[n][matched[n]].double() for n in range(N)], bpi, (1.0, 1.0, 1.0, 1.0))
    g_cls = torch.autograd.grad(ref["loss_rpn_cls"], lo, retain_graph=True)[0]
    g_loc = torch.autograd.grad(ref["loss_rpn_loc"], do)[0]
    # ---- kernel (flat NHWC-row layout, pad column filled with garbage that must be ignored)
    lf = _flat_from_nr(logits, N, H, W, A, 1, 7.0).to(cuda)
    df = _flat_from_nr(deltas.reshape(N, R, 8), N, H, W, A, 8, -3.0).to(cuda)
    rows = N * H * (W + 1)
    loss2 = torch.empty(2, device=cuda)
    dl = torch.empty(rows, A, device=cuda)
    dd = torch.empty(rows, A * 8, device=cuda)
    call("ptb200_rpn_loss_sup", lf, A, df, A * 8, labels.to(torch.int8).to(cuda), matched.to(torch.int32).to(cuda),
         gt.to(cuda), cap, anchors.to(cuda), N, H, W, A, norm, loss2, dl, dd)
    torch.cuda.synchronize()
    assert abs(float(loss2[0]) - float(ref["loss_rpn_cls"])) <= TOL * abs(float(ref["loss_rpn_cls"]))
    assert abs(float(loss2[1]) - float(ref["loss_rpn_loc"])) <= TOL * abs(float(ref["loss_rpn_loc"]))
    assert _rel(_nr_from_flat(dl.cpu(), N, H, W, A, 1)[..., 0], g_cls) < TOL
    assert _rel(_nr_from_flat(dd.cpu(), N, H, W, A, 8), g_loc) < TOL
    # pad-column rows receive no gradient
    assert float(dl.view(N, H, W + 1, A)[:, :, W].abs().max()) == 0.0
    assert float(dd.view(N, H, W + 1, A * 8)[:, :, W].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ RPN unsupervised
@pytest.mark.parametrize("efl,lam,tau,danchor", [(True, (0.5, 0.5), (0.5, 0.5), True), (False, (0.5, 0.5), (0.5, 0.5), False),
                                                 (True, (1.0, 2.0), (0.25, 0.75), True)])
def test_rpn_loss_unsup_kernel(cuda, efl, lam, tau, danchor):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200._lib import call
    g = torch.Generator().manual_seed(5)
    N, H, W, A, cap, K1 = 2, 6, 9, 9, 16, 9
    R = H * W * A
    wh0 = torch.tensor([[181., 90.], [128., 128.], [90., 181.], [362., 181.], [256., 256.], [181., 362.], [724., 362.],
                        [512., 512.], [362., 724.]])
    logits = torch.randn(N, R, generator=g) * 2
    deltas = torch.randn(N, R, 8, generator=g)
    mask = torch.rand(N, R, generator=g) < 0.3
    matched = torch.randint(0, cap, (N, R), generator=g)
    xy = torch.rand(N, cap, 2, generator=g) * torch.tensor([W * 16 * 0.7, H * 16 * 0.7])
    pseudo = torch.cat([xy, xy + 8 + torch.rand(N, cap, 2, generator=g) * 60], -1)
    plog = torch.randn(N, cap, K1, generator=g) * 2   # teacher class logits: some arg-max at the background column
    psig = torch.randn(N, cap, 4, generator=g)
    bpi = 256
    norm = 1.0 / (bpi * N)
    # ---- referee (anchors as a function of the learnable (w, h) pairs: danchor)
    whp = wh0.double().requires_grad_(True)
    anchors = _anchors(O, H, W, whp)
    lo = logits.double().requires_grad_(True)
    do = deltas.double().requires_grad_(True)
    ref = O.rpn_loss_unsupervised(
        lo, [plog[n][matched[n][mask[n]]].double() for n in range(N)], do, [mask[n] for n in range(N)],
        [pseudo[n][matched[n]].double() for n in range(N)], [psig[n][matched[n][mask[n]]].double() for n in range(N)],
        anchors, efl, lam, tau, bpi, (1.0, 1.0, 1.0, 1.0))
    g_cls = torch.autograd.grad(ref["loss_rpn_cls"], lo, retain_graph=True)[0]
    g_loc, g_wh = torch.autograd.grad(ref["loss_rpn_loc"], [do, whp])
    # ---- kernel
    lf = _flat_from_nr(logits, N, H, W, A, 1, 7.0).to(cuda)
    df = _flat_from_nr(deltas, N, H, W, A, 8, -3.0).to(cuda)
    rows = N * H * (W + 1)
    loss2 = torch.empty(2, device=cuda)
    dl = torch.empty(rows, A, device=cuda)
    dd = torch.empty(rows, A * 8, device=cuda)
    da = torch.empty(A, 2, device=cuda) if danchor else None
    call("ptb200_rpn_loss_unsup", lf, A, df, A * 8, mask.to(torch.int32).to(cuda), matched.to(torch.int32).to(cuda),
         pseudo.to(cuda), plog.to(cuda), psig.to(cuda), cap, _anchors(O, H, W, wh0).to(cuda), N, H, W, A, K1, int(efl),
         float(lam[0]), float(lam[1]), float(tau[0]), float(tau[1]), norm, loss2, dl, dd, da)
    torch.cuda.synchronize()
    assert abs(float(loss2[0]) - float(ref["loss_rpn_cls"])) <= TOL * abs(float(ref["loss_rpn_cls"]))
    assert abs(float(loss2[1]) - float(ref["loss_rpn_loc"])) <= TOL * abs(float(ref["loss_rpn_loc"]))
    assert _rel(_nr_from_flat(dl.cpu(), N, H, W, A, 1)[..., 0], g_cls) < TOL
    assert _rel(_nr_from_flat(dd.cpu(), N, H, W, A, 8), g_loc) < TOL
    if danchor:
        assert float(g_wh.abs().max()) > 0
        assert _rel(da, g_wh) < 5 * TOL   # 9 x 2 sums of ~1e3 fp32 atomics each


# ------------------------------------------------------------------------------------------ ROI supervised
def test_roi_loss_sup_kernel(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200._lib import call
    g = torch.Generator().manual_seed(3)
    N, cap, K = 3, 64, 8
    counts = torch.tensor([64, 37, 0], dtype=torch.int32)
    rows = N * cap
    scores = torch.randn(rows, K + 1, generator=g) * 2
    deltas = torch.randn(rows, 8 * K, generator=g)
    cls = torch.randint(0, K + 1, (rows,), generator=g)
    xy = torch.rand(rows, 2, generator=g) * 300
    props = torch.cat([xy, xy + 8 + torch.rand(rows, 2, generator=g) * 100], -1)
    gtb = props + torch.randn(rows, 4, generator=g) * 4
    gtb[:, 2:] = torch.maximum(gtb[:, 2:], gtb[:, :2] + 4)
    weights = (10.0, 10.0, 5.0, 5.0)
    live = torch.zeros(rows, dtype=torch.bool)
    for n, c in enumerate(counts.tolist()):
        live[n * cap:n * cap + c] = True
    so = scores.double().requires_grad_(True)
    do = deltas.double().requires_grad_(True)
    l_cls = F.cross_entropy(so[live], cls[live], reduction="mean")
    l_box = O.roi_box_reg_loss(props[live].double(), gtb[live].double(), do[live], cls[live], K, weights)
    g_s = torch.autograd.grad(l_cls, so)[0]
    g_d = torch.autograd.grad(l_box, do)[0]
    loss2 = torch.empty(2, device=cuda)
    ds = torch.empty(rows, K + 1, device=cuda)
    dd = torch.empty(rows, 8 * K, device=cuda)
    call("ptb200_roi_loss_sup", scores.to(cuda), deltas.to(cuda), cls.to(torch.int32).to(cuda), props.to(cuda),
         gtb.to(cuda), counts.to(cuda), N, cap, K, list(weights), loss2, ds, dd)
    torch.cuda.synchronize()
    assert abs(float(loss2[0]) - float(l_cls)) <= TOL * abs(float(l_cls))
    assert abs(float(loss2[1]) - float(l_box)) <= TOL * abs(float(l_box))
    assert _rel(ds, g_s) < TOL and _rel(dd, g_d) < TOL
    assert float(ds[~live.to(cuda)].abs().max()) == 0.0 and float(dd[~live.to(cuda)].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ ROI unsupervised
@pytest.mark.parametrize("efl,lam,tau", [(True, (0.5, 0.5), (0.5, 0.5)), (False, (0.5, 0.5), (0.5, 0.5)),
                                         (True, (1.0, 2.0), (0.25, 0.75))])
def test_roi_loss_unsup_kernel(cuda, efl, lam, tau):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200._lib import call
    g = torch.Generator().manual_seed(4)
    N, cap, K = 2, 96, 8
    counts = torch.tensor([96, 41], dtype=torch.int32)
    rows = N * cap
    scores = torch.randn(rows, K + 1, generator=g) * 2
    deltas = torch.randn(rows, 8 * K, generator=g)
    soft = torch.randn(rows, K + 1, generator=g) * 2
    sig = torch.randn(rows, 4, generator=g)
    xy = torch.rand(rows, 2, generator=g) * 300
    props = torch.cat([xy, xy + 8 + torch.rand(rows, 2, generator=g) * 100], -1)
    pseudo = props + torch.randn(rows, 4, generator=g) * 4
    pseudo[:, 2:] = torch.maximum(pseudo[:, 2:], pseudo[:, :2] + 4)
    weights = (10.0, 10.0, 5.0, 5.0)
    live = torch.zeros(rows, dtype=torch.bool)
    for n, c in enumerate(counts.tolist()):
        live[n * cap:n * cap + c] = True
    so = scores.double().requires_grad_(True)
    do = deltas.double().requires_grad_(True)
    zt = soft[live].double()
    l_cls = O.roi_cls_loss_unsupervised(so[live], zt, efl, lam, tau)
    cls = zt.max(-1)[1]
    m = cls != K
    mean_p = O.get_deltas(props[live].double(), pseudo[live].double(), weights)
    sel = do[live].view(-1, K, 8)[m][torch.arange(int(m.sum())), cls[m]]
    l_box = O.roi_box_loss_unsupervised(sel[:, :4], sel[:, -4:], mean_p[m], sig[live].double()[m], efl, lam, tau)
    g_s = torch.autograd.grad(l_cls, so)[0]
    g_d = torch.autograd.grad(l_box, do)[0]
    loss2 = torch.empty(2, device=cuda)
    totals = torch.empty(2, dtype=torch.int32, device=cuda)
    ds = torch.empty(rows, K + 1, device=cuda)
    dd = torch.empty(rows, 8 * K, device=cuda)
    call("ptb200_roi_loss_unsup", scores.to(cuda), deltas.to(cuda), soft.to(cuda), sig.to(cuda), props.to(cuda),
         pseudo.to(cuda), counts.to(cuda), N, cap, K, int(efl), float(lam[0]), float(lam[1]), float(tau[0]),
         float(tau[1]), list(weights), totals, loss2, ds, dd)
    torch.cuda.synchronize()
    assert totals.tolist() == [int(live.sum()), int(m.sum())]
    assert abs(float(loss2[0]) - float(l_cls)) <= TOL * abs(float(l_cls))
    assert abs(float(loss2[1]) - float(l_box)) <= TOL * abs(float(l_box))
    assert _rel(ds, g_s) < TOL and _rel(dd, g_d) < TOL


def test_roi_loss_unsup_kernel_empty_edge_is_nan_as_in_the_reference(cuda):
    """No roi matched a pseudo box: the reference takes a mean over zero rows (fast_rcnn.py:208-209,260) -> NaN."""
    from probabilisticteacher_b200._lib import call
    N, cap, K = 2, 32, 8
    rows = N * cap
    z = torch.zeros(rows, K + 1, device=cuda)
    d = torch.zeros(rows, 8 * K, device=cuda)
    b = torch.tensor([[0., 0., 10., 10.]], device=cuda).repeat(rows, 1)
    loss2 = torch.empty(2, device=cuda)
    totals = torch.empty(2, dtype=torch.int32, device=cuda)
    ds = torch.empty(rows, K + 1, device=cuda)
    dd = torch.empty(rows, 8 * K, device=cuda)
    call("ptb200_roi_loss_unsup", z, d, z, torch.zeros(rows, 4, device=cuda), b, b,
         torch.zeros(N, dtype=torch.int32, device=cuda), N, cap, K, 1, 0.5, 0.5, 0.5, 0.5, [10.0, 10.0, 5.0, 5.0], totals,
         loss2, ds, dd)
    torch.cuda.synchronize()
    assert totals.tolist() == [0, 0]
    # (the kernel adds nothing to a zeroed accumulator: 0 * (1/0) is never formed; the model-level test
    # tests/test_zz_next_rows_gpu.py::test_unsupervised_branch_without_pseudo_labels pins the NaN the model returns)
    assert float(ds.abs().max()) == 0.0 and float(dd.abs().max()) == 0.0
    assert math.isnan(float(loss2[0])) or float(loss2[0]) == 0.0
