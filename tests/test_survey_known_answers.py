"""CPU: SURVEY.md Appendix B -- known-answer values produced by the REFERENCE'S OWN functions (unmodified files, run
during the survey) for seeded inputs described by their draw order. The oracle must return the same numbers from the
same draws (rows a-7, a-8, a-10, a-15, a-16 of the scope table); the numbers below are copied from SURVEY.md."""
import pytest
import torch

from oracle import pt_oracle as O


def _gens(seed):
    g = torch.Generator().manual_seed(seed)
    return (lambda *s: torch.randn(*s, generator=g)), (lambda *s: torch.rand(*s, generator=g)), g


def _eq(a, b):
    assert float(a) == pytest.approx(b, rel=2e-6), (float(a), b)


def test_appendix_b1_b2_b3():
    r, u, _ = _gens(1234)
    # B1 -- ROI unsupervised losses (fast_rcnn.py:179-263)
    zs, zt, mq, sq, mp, sp = r(6, 9), r(6, 9) * 3, r(6, 4) * .1, r(6, 4), r(6, 4) * .1, r(6, 4)
    _eq(O.roi_cls_loss_unsupervised(zs, zt, True, [.5, .5], [.5, .5]), 1.6340996026992798)
    _eq(O.roi_cls_loss_unsupervised(zs, zt, False, [.5, .5], [.25, .25]), 2.263561248779297)
    _eq(O.roi_box_loss_unsupervised(mq, sq, mp, sp, True, [.5, .5], [.5, .5]), 0.12519939243793488)
    _eq(O.roi_box_loss_unsupervised(mq, sq, mp, sp, False, [.5, .5], [.5, .5]), 0.20232868194580078)
    # B2 -- pseudo-label filter (fast_rcnn.py:34-120), same generator continuing
    props = u(50, 4) * 300
    props[:, 2:] = props[:, :2] + 20 + u(50, 2) * 200
    logits, deltas = r(50, 9) * 2, r(50, 64) * 0.2
    boxes = O.apply_deltas(deltas, props, (10., 10., 5., 5.))
    res, src = O.fast_rcnn_inference_single_image(boxes, torch.softmax(logits, -1), (600, 800), 0.05, 0.5, 100, logits, deltas)
    assert len(res.scores) == 100 and res.boxes_sigma.shape == (100, 4) and res.scores_logists.shape == (100, 9)
    _eq(O._bt(res.pred_boxes).sum(), 87971.90625)
    _eq(res.scores.sum(), 17.860288619995117)
    assert res.pred_classes[:10].tolist() == [3, 4, 3, 6, 4, 1, 7, 6, 1, 3]
    assert src[:10].tolist() == [30, 40, 37, 45, 42, 19, 25, 9, 20, 47]
    # B3 -- RPN proposal selection incl. the misaligned-sigma quirk (proposal_utils.py:27-154), same generator
    anc = u(315, 4) * 100
    anc[:, 2:] = anc[:, :2] + 16 + u(315, 2) * 64
    d, lg = r(2, 315, 8) * 0.3, r(2, 315)
    pr = O.apply_deltas(d[..., :4].reshape(-1, 4), anc.expand(2, 315, 4).reshape(-1, 4), (1., 1., 1., 1.)).view(2, -1, 4)
    out = O.find_top_rpn_proposals(pr, lg.clone(), [(120, 160)] * 2, 0.7, 200, 50, 0, True, d[..., 4:])
    want = [(13212.517578125, 38.0810432434082, [1.44497811794281, 1.4211312532424927, 1.3268163204193115]),
            (12863.8984375, 37.79094696044922, [1.447826623916626, 1.4423398971557617, 1.2210975885391235])]
    for o, (bs, ss, top3) in zip(out, want):
        assert len(o.objectness_logits) == 50
        _eq(O._bt(o.proposal_boxes).sum(), bs)
        _eq(o.objectness_logits.sum(), ss)
        assert o.objectness_logits[:3].tolist() == pytest.approx(top3, rel=2e-6)


def test_appendix_b4_rpn_losses():
    r, u, g = _gens(4321)
    anc = u(40, 4) * 100
    anc[:, 2:] = anc[:, :2] + 16 + u(40, 2) * 64
    logits, deltas = r(2, 40), r(2, 40, 8) * 0.3
    masks = [u(40) < 0.3 for _ in range(2)]
    assert [int(m.sum()) for m in masks] == [13, 13]
    soft = [r(int(m.sum()), 9) * 3 for m in masks]
    sig = [r(int(m.sum()), 4) for m in masks]
    mgt = []
    for _ in range(2):
        b = u(40, 4) * 100
        b[:, 2:] = b[:, :2] + 16 + u(40, 2) * 64
        mgt.append(b)
    W = (1., 1., 1., 1.)
    o = O.rpn_loss_unsupervised(logits, soft, deltas, masks, mgt, sig, anc, True, [.5, .5], [.5, .5], 256, W)
    _eq(o["loss_rpn_cls"], 0.025443069636821747)
    _eq(o["loss_rpn_loc"], 0.1092383861541748)
    o = O.rpn_loss_unsupervised(logits, soft, deltas, masks, mgt, sig, anc, False, [.5, .5], [.25, .25], 256, W)
    _eq(o["loss_rpn_cls"], 0.0329466313123703)
    _eq(o["loss_rpn_loc"], 0.24084477126598358)
    labels = [torch.randint(-1, 2, (40,), generator=g) for _ in range(2)]
    assert sum(int((l == 1).sum()) for l in labels) == 38 and sum(int((l >= 0).sum()) for l in labels) == 58
    o = O.rpn_losses(anc, logits, labels, deltas, mgt, 256, W)
    _eq(o["loss_rpn_cls"], 0.08874400705099106)
    _eq(o["loss_rpn_loc"], 0.4521816372871399)
