"""GPU parity of the non-GEMM kernels (through the C ABI) against the CPU oracle on the same seeded
inputs. Integer / index outputs must be bit-exact; floating point within the stated tolerance."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _boxes(g, n, W=600, H=400, lo=8, hi=200):
    xy = torch.rand(n, 2, generator=g) * torch.tensor([W - lo, H - lo])
    wh = torch.rand(n, 2, generator=g) * (hi - lo) + lo
    return torch.cat([xy, xy + wh], 1)


def test_segmented_sort(cuda):
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(0)
    segs, stride = 5, 40000
    lens = torch.tensor([40000, 1, 0, 12345, 4096], dtype=torch.int32)
    keys = torch.randint(0, 2 ** 31 - 1, (segs, stride), generator=g, dtype=torch.int64)
    keys[3, :5000] = keys[3, 0]  # duplicates: stability
    keys[0] = keys[0] % 1000
    vals = torch.arange(stride, dtype=torch.int32).repeat(segs, 1)
    k = keys.to(torch.int32).to(cuda)
    v = vals.to(cuda)
    ops.segmented_sort(k, v, seg_len=lens.to(cuda))
    torch.cuda.synchronize()
    for s in range(segs):
        n = int(lens[s])
        rk, ri = torch.sort(keys[s, :n], stable=True)
        assert torch.equal(k[s, :n].cpu().to(torch.int64), rk)
        assert torch.equal(v[s, :n].cpu().to(torch.int64), ri)


@pytest.mark.parametrize("n,thr,classes", [(300, 0.7, 0), (5000, 0.7, 0), (12000, 0.7, 0), (4000, 0.5, 8)])
def test_nms_bit_exact(cuda, n, thr, classes):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(n)
    centers = _boxes(g, max(n // 20, 4))
    b = centers[torch.randint(0, centers.shape[0], (n,), generator=g)] + torch.randn(n, 4, generator=g) * 6
    b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 1)
    s = torch.rand(n, generator=g)
    order = torch.argsort(-s, stable=True).to(torch.int32)
    if classes:
        cls = (torch.arange(n) % classes)
        ref = O.batched_nms(b, s, cls, thr)
    else:
        ref = O.nms(b, s, thr)
    max_keep = 2000
    cap = ((n + 63) // 64) * 64
    order_p = torch.zeros(1, cap, dtype=torch.int32)
    order_p[0, :n] = order
    keep_idx, keep_count = ops.nms(b[None].to(cuda), order_p.to(cuda), torch.tensor([n], dtype=torch.int32).to(cuda),
                                   thr, max_keep, class_mod=classes)
    torch.cuda.synchronize()
    kc = int(keep_count[0])
    got = order[keep_idx[0, :kc].cpu().long()].long()
    assert kc == min(len(ref), max_keep)
    assert torch.equal(got, ref[:max_keep])


def test_oracle_nms_matches_torchvision():
    pass  # covered in the CPU suite (tests/test_oracle_cpu.py)


@pytest.mark.parametrize("M", [0, 1, 13, 60])
def test_rpn_match_and_subsample(cuda, M):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.modeling import sampling
    g = torch.Generator().manual_seed(100 + M)
    H, W = 25, 38
    cell = O.default_cell_anchors((64, 128, 256), (0.5, 1.0, 2.0))
    anchors = O.grid_anchors(cell, H, W, 16, 0.0)
    R = anchors.shape[0]
    N, cap = 2, 64
    gt = torch.zeros(N, cap, 4)
    cnt = torch.tensor([M, max(M - 1, 0)], dtype=torch.int32)
    for n in range(N):
        gt[n, :cnt[n]] = _boxes(g, int(cnt[n]))
    matched, labels = sampling.rpn_match(gt.to(cuda), cnt.to(cuda), anchors.to(cuda), N, 0.3, 0.7)
    prio_p = torch.rand(N, R, generator=g)
    prio_n = torch.rand(N, R, generator=g)
    sampled = sampling.rpn_subsample(labels, 256, 0.25, prio_p.to(cuda), prio_n.to(cuda))
    torch.cuda.synchronize()
    for n in range(N):
        iou = O.pairwise_iou(gt[n, :cnt[n]], anchors)
        rm, rl = O.matcher(iou, (0.3, 0.7), (0, -1, 1), True)
        assert torch.equal(labels[n].cpu().to(torch.int8), rl)
        if cnt[n] > 0:
            assert torch.equal(matched[n].cpu().long(), rm)
        pos, neg = O.subsample_labels(rl, 256, 0.25, 0, prio_p[n], prio_n[n])
        ref = torch.full((R,), -1, dtype=torch.int8)
        ref[pos] = 1
        ref[neg] = 0
        assert torch.equal(sampled[n].cpu(), ref)


def test_roi_label_and_sample(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.modeling import sampling
    g = torch.Generator().manual_seed(5)
    N, pcap, gcap, K = 2, 300, 16, 8
    gtc = torch.tensor([12, 0], dtype=torch.int32)
    pc = torch.tensor([300, 257], dtype=torch.int32)
    gt = torch.zeros(N, gcap, 4)
    gcls = torch.zeros(N, gcap, dtype=torch.int32)
    props = torch.zeros(N, pcap, 4)
    for n in range(N):
        gt[n, :gtc[n]] = _boxes(g, int(gtc[n]))
        gcls[n, :gtc[n]] = torch.randint(0, K, (int(gtc[n]),), generator=g, dtype=torch.int32)
        p = _boxes(g, int(pc[n]))
        if gtc[n] > 0:  # some proposals near gt
            idx = torch.randint(0, int(gtc[n]), (100,), generator=g)
            p[:100] = gt[n, idx] + torch.randn(100, 4, generator=g) * 4
        props[n, :pc[n]] = p
    L = pcap + gcap
    prio_p = torch.rand(N, L, generator=g)
    prio_n = torch.rand(N, L, generator=g)
    out = sampling.roi_label_and_sample(gt.to(cuda), gcls.to(cuda), gtc.to(cuda), props.to(cuda), pc.to(cuda), K, 0.5,
                                        128, 0.25, prio_p.to(cuda), prio_n.to(cuda))
    torch.cuda.synchronize()
    for n in range(N):
        gtb = gt[n, :gtc[n]]
        pb = torch.cat([props[n, :pc[n]], gtb], 0)
        iou = O.pairwise_iou(gtb, pb)
        midx, mlab = O.matcher(iou, [0.5], [0, 1], False)
        if gtc[n] > 0:
            cls = gcls[n, :gtc[n]].long()[midx].clone()
            cls[mlab == 0] = K
        else:
            cls = torch.zeros_like(midx) + K
        fg, bg = O.subsample_labels(cls, 128, 0.25, K, prio_p[n], prio_n[n])
        sidx = torch.cat([fg, bg])
        c = int(out["count"][n])
        assert c == sidx.numel()
        assert torch.equal(out["src"][n, :c].cpu().long(), sidx)
        assert torch.equal(out["gt_classes"][n, :c].cpu().long(), cls[sidx])
        assert torch.equal(out["rois"][n, :c].cpu(), pb[sidx])
        if gtc[n] > 0:
            assert torch.equal(out["gt_boxes"][n, :c].cpu(), gtb[midx[sidx]])


def test_roi_match_unsup(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.modeling import sampling
    g = torch.Generator().manual_seed(6)
    N, pcap, scap, K1 = 2, 1500, 100, 9
    sc = torch.tensor([37, 100], dtype=torch.int32)
    pc = torch.tensor([1500, 1111], dtype=torch.int32)
    ps = torch.zeros(N, scap, 4)
    pl = torch.randn(N, scap, K1, generator=g)
    sg = torch.randn(N, scap, 4, generator=g)
    props = torch.zeros(N, pcap, 4)
    for n in range(N):
        ps[n, :sc[n]] = _boxes(g, int(sc[n]))
        idx = torch.randint(0, int(sc[n]), (int(pc[n]),), generator=g)
        props[n, :pc[n]] = ps[n, idx] + torch.randn(int(pc[n]), 4, generator=g) * 12
        props[n, :pc[n], 2:] = torch.maximum(props[n, :pc[n], 2:], props[n, :pc[n], :2] + 1)
    out = sampling.roi_match_unsup(ps.to(cuda), pl.to(cuda), sg.to(cuda), sc.to(cuda), props.to(cuda), pc.to(cuda), 0.5)
    torch.cuda.synchronize()
    for n in range(N):
        iou = O.pairwise_iou(ps[n, :sc[n]], props[n, :pc[n]])
        midx, mlab = O.matcher(iou, [0.5], [0, 1], False)
        sel = mlab == 1
        c = int(out["count"][n])
        assert c == int(sel.sum())
        assert torch.equal(out["rois"][n, :c].cpu(), props[n, :pc[n]][sel])
        assert torch.equal(out["pseudo_boxes"][n, :c].cpu(), ps[n][midx][sel])
        assert torch.equal(out["soft_label"][n, :c].cpu(), pl[n][midx][sel])
        assert torch.equal(out["boxes_sigma"][n, :c].cpu(), sg[n][midx][sel])


def test_maxpool_fwd_bwd(cuda):
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(8)
    x = torch.relu(torch.randn(2, 64, 21, 35, generator=g)).half().float().requires_grad_(True)
    y = torch.nn.functional.max_pool2d(x, 2, 2)
    gy = torch.randn(y.shape, generator=g).half().float()
    y.backward(gy)
    ref_dx = x.grad * (x > 0)
    xa = ops.to_flat(x.detach().to(cuda))
    ya = ops.maxpool2x2(xa)
    dz = ops.maxpool2x2_relu_bwd(xa, ops.to_flat(gy.to(cuda)))
    torch.cuda.synchronize()
    assert torch.equal(ops.from_flat(ya).float().cpu(), y.detach())
    assert ya.t.reshape(2, ya.H, ya.W + 1, 64)[:, :, ya.W].abs().max() == 0
    assert torch.equal(ops.from_flat(dz).float().cpu(), ref_dx)


def test_roi_align_fwd_bwd(cuda):
    from torchvision.ops import roi_align
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(9)
    N, C, H, W, cap = 2, 64, 19, 30, 40
    feat = torch.randn(N, C, H, W, generator=g).half().float().requires_grad_(True)
    rois = torch.zeros(N, cap, 4)
    cnt = torch.tensor([40, 23], dtype=torch.int32)
    for n in range(N):
        rois[n] = _boxes(g, cap, W=W * 16, H=H * 16, lo=4, hi=300)
    rois[0, 0] = torch.tensor([-30.0, -20.0, 700.0, 500.0])  # partly outside
    rl = torch.cat([torch.cat([torch.full((int(cnt[n]), 1), float(n)), rois[n, :cnt[n]]], 1) for n in range(N)])
    ref = roi_align(feat, rl, (7, 7), 1.0 / 16, 0, True)
    gout = torch.randn(ref.shape, generator=g).half().float()
    ref.backward(gout)
    fa = ops.to_flat(feat.detach().to(cuda))
    out = ops.roi_align_fwd(fa, rois.to(cuda), cnt.to(cuda), cap, 1.0 / 16, 7)
    out = out.reshape(N, cap, 49, C)
    got = torch.cat([out[n, :cnt[n]] for n in range(N)]).permute(0, 2, 1).reshape(-1, C, 7, 7).float().cpu()
    assert (got - ref.detach()).abs().max() < 2e-3 * ref.abs().max()
    dout = torch.zeros(N, cap, 49, C, dtype=torch.float16)
    o = 0
    for n in range(N):
        c = int(cnt[n])
        dout[n, :c] = gout[o:o + c].reshape(c, C, 49).permute(0, 2, 1).half()
        o += c
    df = ops.roi_align_bwd(dout.reshape(N * cap, -1).to(cuda), fa, rois.to(cuda), cnt.to(cuda), cap, 1.0 / 16, 7)
    torch.cuda.synchronize()
    dfm = df.reshape(N, H, W + 1, C)[:, :, :W].permute(0, 3, 1, 2).cpu()
    assert (dfm - feat.grad).abs().max() < 1e-3 * feat.grad.abs().max()


def test_preprocess_im2col(cuda):
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(10)
    N, H, W = 2, 13, 22
    img = torch.randint(0, 256, (N, 3, H, W), generator=g, dtype=torch.uint8)
    mean, std = (103.53, 116.28, 123.675), (1.0, 1.0, 1.0)
    hw = torch.tensor([[H, W]] * N, dtype=torch.int32)
    a = ops.preprocess_im2col(img.to(cuda), hw.to(cuda), H, W, mean, std)
    torch.cuda.synchronize()
    x = (img.float() - torch.tensor(mean).view(1, 3, 1, 1))
    cols = torch.nn.functional.unfold(x, 3, padding=1).reshape(N, 3, 9, H, W)  # [N, c, t, H, W]
    ref = cols.permute(0, 3, 4, 2, 1).reshape(N, H, W, 27).half()
    got = a.t.reshape(N, H, W + 1, 64)
    assert torch.equal(got[:, :, :W, :27].cpu(), ref)
    assert got[:, :, :, 27:].abs().max() == 0 and got[:, :, W].abs().max() == 0


def test_find_top_rpn_proposals(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.modeling.proposal_generator.proposal_utils import find_top_rpn_proposals
    g = torch.Generator().manual_seed(11)
    N, H, W, A = 2, 20, 31, 9
    R = H * W * A
    cell = O.default_cell_anchors((64, 128, 256), (0.5, 1.0, 2.0))
    anchors = O.grid_anchors(cell, H, W, 16, 0.0)
    logits = torch.randn(N, R, generator=g)
    deltas = torch.randn(N, R, 8, generator=g) * 0.3
    sizes = [(H * 16, W * 16), (H * 16 - 30, W * 16 - 50)]
    props = O.apply_deltas(deltas[..., :4].reshape(-1, 4), anchors.unsqueeze(0).expand(N, -1, -1).reshape(-1, 4),
                           (1, 1, 1, 1)).view(N, -1, 4)
    ref = O.find_top_rpn_proposals(props, logits, sizes, 0.7, 3000, 500, 0.0, True, deltas[..., 4:])
    # device layout: flat rows with pad column
    lg = torch.zeros(N, H, W + 1, A)
    lg[:, :, :W] = logits.view(N, H, W, A)
    dl = torch.zeros(N, H, W + 1, A * 8)
    dl[:, :, :W] = deltas.view(N, H, W, A * 8)
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    hw = torch.tensor(sizes, dtype=torch.float32)
    boxes, scores, count = find_top_rpn_proposals(lg.reshape(N, -1, A).to(cuda), dl.reshape(N, -1, A * 8).to(cuda),
                                                  anchors.to(cuda), N, H, W, A, hw.to(cuda), 0.7, 3000, 500, 0.0, flag)
    torch.cuda.synchronize()
    assert int(flag) == 0
    for n in range(N):
        c = int(count[n])
        rb = ref[n].proposal_boxes.tensor
        assert c == rb.shape[0]
        assert (boxes[n, :c].cpu() - rb).abs().max() < 1e-2
        assert (scores[n, :c].cpu() - ref[n].objectness_logits).abs().max() < 1e-5


def test_roi_inference_filter(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200._lib import call
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(12)
    N, cap, K = 2, 200, 8
    cnt = torch.tensor([200, 150], dtype=torch.int32)
    props = torch.stack([_boxes(g, cap) for _ in range(N)])
    scores = torch.randn(N, cap, K + 1, generator=g) * 2
    deltas = torch.randn(N, cap, 8 * K, generator=g) * 0.2
    sizes = [(400, 600), (380, 560)]
    w = (10.0, 10.0, 5.0, 5.0)
    dev = cuda
    cb = torch.empty(N, cap * K, 4, device=dev)
    cs = torch.empty(N, cap * K, device=dev)
    keys = torch.empty(N, cap * K, dtype=torch.int32, device=dev)
    vals = torch.empty(N, cap * K, dtype=torch.int32, device=dev)
    cc = torch.empty(N, dtype=torch.int32, device=dev)
    hw = torch.tensor(sizes, dtype=torch.float32, device=dev)
    sd, dd = scores.to(dev), deltas.to(dev)
    call("ptb200_roi_infer_candidates", sd, dd, props.to(dev), cnt.to(dev), N, cap, K, hw, 0.05, list(w), cb, cs, keys,
         vals, cc)
    ops.segmented_sort(keys, vals)
    keep_idx, keep_count = ops.nms(cb, vals, cc, 0.5, 100, class_mod=K)
    ob = torch.empty(N, 100, 4, device=dev)
    osc = torch.empty(N, 100, device=dev)
    ocl = torch.empty(N, 100, dtype=torch.int64, device=dev)
    olg = torch.empty(N, 100, K + 1, device=dev)
    osg = torch.empty(N, 100, 4, device=dev)
    osr = torch.empty(N, 100, dtype=torch.int32, device=dev)
    call("ptb200_roi_infer_gather", cb, cs, sd, dd, vals, keep_idx, keep_count, N, cap, K, 100, ob, osc, ocl, olg, osg,
         osr)
    torch.cuda.synchronize()
    for n in range(N):
        c = int(cnt[n])
        b = O.apply_deltas(deltas[n, :c], props[n, :c], w)
        r, src = O.fast_rcnn_inference_single_image(b, torch.softmax(scores[n, :c], -1), sizes[n], 0.05, 0.5, 100,
                                                    scores[n, :c], deltas[n, :c])
        kc = int(keep_count[n])
        assert kc == len(r.scores)
        assert torch.equal(osr[n, :kc].cpu().long(), src)
        assert torch.equal(ocl[n, :kc].cpu(), r.pred_classes)
        assert (ob[n, :kc].cpu() - r.pred_boxes.tensor).abs().max() < 1e-2
        assert (osc[n, :kc].cpu() - r.scores).abs().max() < 1e-5
        assert torch.equal(olg[n, :kc].cpu(), r.scores_logists)
        assert torch.equal(osg[n, :kc].cpu(), r.boxes_sigma)


def test_ema_and_sgd(cuda):
    from probabilisticteacher_b200._lib import call
    g = torch.Generator().manual_seed(13)
    n = 100003
    t = torch.randn(n, generator=g)
    s = torch.randn(n, generator=g)
    td = t.to(cuda)
    call("ptb200_ema_update", td, s.to(cuda), n, 0.9996)
    assert torch.allclose(td.cpu(), s * (1 - 0.9996) + t * 0.9996, atol=1e-6)
    p = torch.randn(n, generator=g)
    gr = torch.randn(n, generator=g) * 3
    m = torch.randn(n, generator=g)
    pp = torch.nn.Parameter(p.clone())
    opt = torch.optim.SGD([pp], lr=0.016, momentum=0.9, weight_decay=1e-4)
    opt.state[pp]["momentum_buffer"] = m.clone()
    norm = gr.norm()
    pp.grad = gr * (10.0 / max(float(norm), 10.0))
    opt.step()
    pd, gd, md = p.to(cuda), gr.to(cuda), m.to(cuda)
    ss = torch.zeros(1, device=cuda)
    call("ptb200_grad_sumsq", gd, n, 1.0, ss)
    call("ptb200_clip_sgd_step", pd, gd, md, n, 0.016, 0.9, 1e-4, 10.0, 1.0, ss)
    torch.cuda.synchronize()
    assert abs(float(ss.sqrt()) - float(norm)) < 1e-2
    assert torch.allclose(pd.cpu(), pp.detach(), atol=1e-5)


def test_conv1_fused_preprocess(cuda):
    from probabilisticteacher_b200 import ops
    from probabilisticteacher_b200._lib import call
    g = torch.Generator().manual_seed(14)
    N, H, W = 2, 37, 53
    img = torch.randint(0, 256, (N, 3, H, W), generator=g, dtype=torch.uint8)
    w = torch.randn(64, 3, 3, 3, generator=g) * 0.05
    b = torch.randn(64, generator=g)
    mean, std = (103.53, 116.28, 123.675), (1.0, 1.0, 1.0)
    wi = w.permute(0, 2, 3, 1).reshape(64, 27).contiguous().to(cuda)
    wp = torch.zeros(64, 32, dtype=torch.float16, device=cuda)
    call("ptb200_cast_pad_rows_f16", wi, wp, 64, 27, 32)
    hw = torch.tensor([[H, W]] * N, dtype=torch.int32).to(cuda)
    a = ops.conv1_u8(img.to(cuda).view(N, -1), hw, H, W, mean, std, wp, b.to(cuda))
    torch.cuda.synchronize()
    x = (img.float() - torch.tensor(mean).view(1, 3, 1, 1)).half().float()
    ref = torch.relu(torch.nn.functional.conv2d(x, w.half().float(), b, padding=1))
    got = a.t.reshape(N, H, W + 1, 64)
    assert got[:, :, W].abs().max() == 0
    r = got[:, :, :W].permute(0, 3, 1, 2).float().cpu()
    assert (r - ref).abs().max() < 2e-3 * ref.abs().max()
