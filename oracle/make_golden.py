"""Generates tests/golden/pt_reference_golden.pt by executing the REFERENCE's own hot-path code
(imported unmodified from /root/reference through oracle/d2shim.py) on seeded inputs.
Run here (the reference tree does not exist on the GPU box):

    python oracle/make_golden.py

The fixture stores inputs and outputs; tests/test_oracle_golden.py checks oracle/pt_oracle.py against
it, which pins the restatement to the reference for every function listed in SURVEY.md 8a that lives
in the reference tree (box transform, Gaussian pdf, RPN / ROI losses, proposal selection with the
misaligned-sigma quirk, pseudo-label filter, anchor labelling, unsup ROI sampling, anchors, VGG block).
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import d2shim  # noqa: E402

d2shim.install()

from detectron2.structures import Boxes  # noqa: E402
from pt.modeling import box_regression as ref_br  # noqa: E402
from pt.modeling.proposal_generator import proposal_utils as ref_pu  # noqa: E402
from pt.modeling.proposal_generator import rpn as ref_rpn  # noqa: E402
from pt.modeling.roi_heads import fast_rcnn as ref_fr  # noqa: E402
from pt.modeling.roi_heads import roi_heads as ref_rh  # noqa: E402
from pt.modeling import anchor_generator as ref_ag  # noqa: E402
from pt.modeling.backbone import vgg as ref_vgg  # noqa: E402
from pt.structures.instances import FreeInstances  # noqa: E402
from detectron2.modeling.matcher import Matcher  # noqa: E402

NS = types.SimpleNamespace


def rand_boxes(g, n, W=600.0, H=400.0, lo=8.0, hi=200.0):
    xy = torch.rand(n, 2, generator=g) * torch.tensor([W - lo, H - lo])
    wh = torch.rand(n, 2, generator=g) * (hi - lo) + lo
    return torch.cat([xy, xy + wh], 1)


def main():
    out = {}
    g = torch.Generator().manual_seed(1234)
    cfg = NS(UNSUPNET=NS(MODEL_TYPE="GUASSIAN"))

    # ---- G1 box transform
    src, tgt = rand_boxes(g, 64), rand_boxes(g, 64)
    deltas = torch.randn(64, 32, generator=g) * 0.3
    case = dict(src=src, tgt=tgt, deltas=deltas)
    for w in ((1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)):
        t = ref_br.Box2BoxTransform(weights=w)
        case[f"get_deltas_{int(w[0])}"] = t.get_deltas(src, tgt)
        case[f"apply_deltas_{int(w[0])}"] = t.apply_deltas(deltas, src)
    val, mean, var = torch.randn(50, 4, generator=g), torch.randn(50, 4, generator=g), torch.rand(50, 4, generator=g)
    case.update(pdf_val=val, pdf_mean=mean, pdf_var=var, pdf=ref_br.gaussian_dist_pdf(val, mean, var))
    out["box_transform"] = case

    # ---- G2 ROI unsupervised losses
    n, K = 40, 8
    zs, zt = torch.randn(n, K + 1, generator=g), torch.randn(n, K + 1, generator=g) * 3
    mq, sq = torch.randn(n, 4, generator=g) * 0.1, torch.randn(n, 4, generator=g)
    mp, sp = torch.randn(n, 4, generator=g) * 0.1, torch.randn(n, 4, generator=g)
    me = NS(model_type="GUASSIAN")
    F_ = ref_fr.GuassianFastRCNNOutputLayers
    case = dict(zs=zs, zt=zt, mq=mq, sq=sq, mp=mp, sp=sp)
    for efl in (True, False):
        for tau in ((0.5, 0.5), (0.25, 0.25)):
            key = f"efl{int(efl)}_tau{tau[0]}"
            case["cls_" + key] = F_.cls_loss_unsupervised(me, zs, zt, efl, [0.5, 0.5], list(tau))["loss_cls"]
            case["box_" + key] = F_.box_reg_loss_unsupervised(me, mq, sq, mp, sp, efl, [0.5, 0.5], list(tau))["loss_box_reg"]
    out["roi_unsup_losses"] = case

    # ---- G3 ROI supervised box regression loss
    R = 64
    props, gts = rand_boxes(g, R), rand_boxes(g, R)
    pd = torch.randn(R, K * 8, generator=g) * 0.2
    gc = torch.randint(0, K + 1, (R,), generator=g)
    me = NS(model_type="GUASSIAN", num_classes=K, box2box_transform=ref_br.Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0)))
    out["roi_box_reg_loss"] = dict(props=props, gts=gts, pred_deltas=pd, gt_classes=gc,
                                   loss=F_.box_reg_loss(me, props, gts, pd, gc))

    # ---- G4 pseudo-label filter
    R = 120
    props = rand_boxes(g, R)
    logits = torch.randn(R, K + 1, generator=g) * 2
    deltas = torch.randn(R, K * 8, generator=g) * 0.2
    boxes = ref_br.Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0)).apply_deltas(deltas, props)
    scores = torch.softmax(logits, -1)
    res, src_idx = ref_fr.fast_rcnn_inference_single_image(boxes, scores, (400, 600), 0.05, 0.5, 100, logits, deltas)
    out["roi_inference"] = dict(props=props, logits=logits, deltas=deltas, image_shape=(400, 600),
                                pred_boxes=res.pred_boxes.tensor, scores=res.scores, pred_classes=res.pred_classes,
                                scores_logists=res.scores_logists, boxes_sigma=res.boxes_sigma, src_idx=src_idx)

    # ---- G5 RPN proposal selection (misaligned-sigma quirk, proposal_utils.py:94)
    N, Rr = 2, 5 * 7 * 9
    anc = rand_boxes(g, Rr, W=160.0, H=120.0, lo=16.0, hi=80.0)
    d = torch.randn(N, Rr, 8, generator=g) * 0.3
    lg = torch.randn(N, Rr, generator=g)
    pr = ref_br.Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0)).apply_deltas(
        d[..., :4].reshape(-1, 4), anc.unsqueeze(0).expand(N, -1, -1).reshape(-1, 4)).view(N, -1, 4)
    res = ref_pu.find_top_rpn_proposals([pr], [lg.clone()], [(120, 160)] * N, 0.7, 200, 50, 0, True, [d[..., 4:]])
    out["rpn_proposals"] = dict(anchors=anc, deltas=d, logits=lg, proposals=pr, image_sizes=[(120, 160)] * N,
                                boxes=[r.proposal_boxes.tensor for r in res],
                                scores=[r.objectness_logits for r in res])

    # ---- G6 RPN losses
    g2 = torch.Generator().manual_seed(4321)
    N, Rr = 2, 60
    anc = rand_boxes(g2, Rr, W=160.0, H=120.0, lo=16.0, hi=80.0)
    logits = torch.randn(N, Rr, generator=g2)
    deltas = torch.randn(N, Rr, 8, generator=g2) * 0.3
    masks = [torch.rand(Rr, generator=g2) < 0.3 for _ in range(N)]
    soft = [torch.randn(int(m.sum()), K + 1, generator=g2) * 3 for m in masks]
    sig = [torch.randn(int(m.sum()), 4, generator=g2) for m in masks]
    mgt = [rand_boxes(g2, Rr, W=160.0, H=120.0, lo=16.0, hi=80.0) for _ in range(N)]
    me = NS(cfg=cfg, batch_size_per_image=256, box2box_transform=ref_br.Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0)),
            box_reg_loss_type="smooth_l1", smooth_l1_beta=0.0, loss_weight={})
    case = dict(anchors=anc, logits=logits, deltas=deltas, masks=masks, soft=soft, sig=sig, mgt=mgt)
    for efl in (True, False):
        for tau in ((0.5, 0.5), (0.25, 0.25)):
            r = ref_rpn.GuassianRPN.loss_rpn_unsupervised(me, [logits], soft, [deltas], masks, mgt, sig, [Boxes(anc)],
                                                          efl, [0.5, 0.5], list(tau), True)
            case[f"unsup_efl{int(efl)}_tau{tau[0]}"] = (r["loss_rpn_cls"], r["loss_rpn_loc"])
    labels = [torch.randint(-1, 2, (Rr,), generator=g2) for _ in range(N)]
    r = ref_rpn.GuassianRPN.losses(me, [Boxes(anc)], [logits], labels, [deltas], mgt)
    case.update(labels=labels, sup=(r["loss_rpn_cls"], r["loss_rpn_loc"]))
    out["rpn_losses"] = case

    # ---- G7 label_and_sample_anchors (supervised with injected priorities; unsupervised soft labels)
    from oracle import pt_oracle as O
    H, W = 12, 17
    cell = O.default_cell_anchors((64, 128, 256), (0.5, 1.0, 2.0))
    anchors = O.grid_anchors(cell, H, W, 16, 0.0)
    Rr = anchors.shape[0]
    gtb = [rand_boxes(g, 7, W=272.0, H=192.0, lo=20.0, hi=150.0), rand_boxes(g, 3, W=272.0, H=192.0, lo=20.0, hi=150.0)]
    prio = [(torch.rand(Rr, generator=g), torch.rand(Rr, generator=g)) for _ in range(2)]
    state = {"i": 0}

    def _subsample(label):
        pp, pn = prio[state["i"]]
        state["i"] += 1
        pos, neg = O.subsample_labels(label, 256, 0.25, 0, pp, pn)
        label.fill_(-1)
        label.scatter_(0, pos, 1)
        label.scatter_(0, neg, 0)
        return label
    me = NS(anchor_matcher=Matcher([0.3, 0.7], [0, -1, 1], True), anchor_boundary_thresh=-1, _subsample_labels=_subsample)
    insts = [FreeInstances((192, 272), gt_boxes=Boxes(b)) for b in gtb]
    lab, mg = ref_rpn.GuassianRPN.label_and_sample_anchors.__wrapped__(me, [Boxes(anchors)], insts) \
        if hasattr(ref_rpn.GuassianRPN.label_and_sample_anchors, "__wrapped__") else \
        ref_rpn.GuassianRPN.label_and_sample_anchors(me, [Boxes(anchors)], insts)
    case = dict(anchors=anchors, gt_boxes=gtb, prio=prio, labels=lab, matched_gt=mg)
    ps = [rand_boxes(g, 9, W=272.0, H=192.0, lo=20.0, hi=150.0), rand_boxes(g, 5, W=272.0, H=192.0, lo=20.0, hi=150.0)]
    pl = [torch.randn(b.shape[0], K + 1, generator=g) for b in ps]
    psg = [torch.randn(b.shape[0], 4, generator=g) for b in ps]
    insts = [FreeInstances((192, 272), pseudo_boxes=Boxes(b), scores_logists=l, boxes_sigma=s) for b, l, s in zip(ps, pl, psg)]
    gl, am, mg2, ms = ref_rpn.GuassianRPN.label_and_sample_anchors(me, [Boxes(anchors)], insts, use_ignore=True,
                                                                   use_soft_label=True)
    case.update(pseudo_boxes=ps, pseudo_logits=pl, pseudo_sigma=psg, u_labels=gl, u_masks=am, u_matched=mg2, u_sigma=ms)
    out["rpn_labelling"] = case

    # ---- G8 DifferentiableAnchorGenerator
    torch.Tensor.cuda = lambda self, *a, **k: self
    wh = [[181.0193, 90.5097], [128.0, 128.0], [90.5097, 181.0193], [362.0387, 181.0193], [256.0, 256.0],
          [181.0193, 362.0387], [724.0773, 362.0387], [512.0, 512.0], [362.0387, 724.0773]]
    gen = ref_ag.DifferentiableAnchorGenerator(anchor=[wh], strides=[16], offset=0.0)
    a = gen([torch.zeros(1, 8, 5, 7)])[0].tensor
    out["anchors"] = dict(wh=torch.tensor(wh), H=5, W=7, anchors=a.detach())

    # ---- G9 unsupervised ROI sampling (roi_heads.py:257-291)
    N = 2
    props = [rand_boxes(g, 80), rand_boxes(g, 60)]
    for k in range(N):
        idx = torch.randint(0, ps[k].shape[0], (30,), generator=g)
        props[k][:30] = ps[k][idx] + torch.randn(30, 4, generator=g) * 5
    me = NS(proposal_append_gt=True, proposal_matcher=Matcher([0.5], [0, 1], False), num_classes=K,
            _sample_proposals_unsup=lambda *a: ref_rh.GuassianROIHead._sample_proposals_unsup(None, *a))
    pin = [FreeInstances((192, 272), proposal_boxes=Boxes(p.clone()), objectness_logits=torch.zeros(p.shape[0])) for p in props]
    res = ref_rh.GuassianROIHead.label_and_sample_proposals(me, pin, insts, branch="unsupervised")
    out["roi_unsup_sampling"] = dict(props=props, boxes=[r.proposal_boxes.tensor for r in res],
                                     pseudo=[r.pseudo_boxes.tensor for r in res], soft=[r.soft_label for r in res],
                                     sigma=[r.boxes_sigma for r in res])

    # ---- G10 VGG block (conv + bias + ReLU x2, 2x2 max pool)
    torch.manual_seed(5)
    blk = ref_vgg.VGGBlock(3, [8, 8], norm="None", pool=True)
    x = torch.randn(1, 3, 11, 14, generator=g)
    with torch.no_grad():
        for c in blk.convs:
            c.bias.normal_(0, 0.1)
        y = blk(x.clone())
    out["vgg_block"] = dict(x=x, w=[c.weight.detach().clone() for c in blk.convs],
                            b=[c.bias.detach().clone() for c in blk.convs], y=y)

    # ---- G11 add_ground_truth_to_proposals
    p = FreeInstances((192, 272), proposal_boxes=Boxes(props[0].clone()), objectness_logits=torch.randn(80, generator=g))
    r = ref_pu.add_ground_truth_to_proposals([Boxes(gtb[0])], [p])[0]
    out["append_gt"] = dict(gt=gtb[0], props=props[0], boxes=r.proposal_boxes.tensor, logits=r.objectness_logits)

    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
