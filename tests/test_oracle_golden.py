"""CPU: the oracle restatement (oracle/pt_oracle.py) against golden vectors produced by the
reference's own code (oracle/make_golden.py -> tests/golden/pt_reference_golden.pt)."""
import os

import pytest
import torch

from oracle import pt_oracle as O

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_golden.pt"), weights_only=False)


def close(a, b, tol=1e-6):
    a, b = torch.as_tensor(a).float(), torch.as_tensor(b).float()
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.allclose(a, b, rtol=tol, atol=tol), float((a - b).abs().max())


def test_box_transform():
    c = G["box_transform"]
    for w in ((1.0, 1.0, 1.0, 1.0), (10.0, 10.0, 5.0, 5.0)):
        close(O.get_deltas(c["src"], c["tgt"], w), c[f"get_deltas_{int(w[0])}"])
        close(O.apply_deltas(c["deltas"], c["src"], w), c[f"apply_deltas_{int(w[0])}"], 1e-5)
    close(O.gaussian_dist_pdf(c["pdf_val"], c["pdf_mean"], c["pdf_var"]), c["pdf"])


def test_roi_unsup_losses():
    c = G["roi_unsup_losses"]
    for efl in (True, False):
        for tau in ((0.5, 0.5), (0.25, 0.25)):
            key = f"efl{int(efl)}_tau{tau[0]}"
            close(O.roi_cls_loss_unsupervised(c["zs"], c["zt"], efl, (0.5, 0.5), tau), c["cls_" + key])
            close(O.roi_box_loss_unsupervised(c["mq"], c["sq"], c["mp"], c["sp"], efl, (0.5, 0.5), tau), c["box_" + key])


def test_roi_box_reg_loss():
    c = G["roi_box_reg_loss"]
    close(O.roi_box_reg_loss(c["props"], c["gts"], c["pred_deltas"], c["gt_classes"], 8, (10.0, 10.0, 5.0, 5.0)), c["loss"])


def test_pseudo_label_filter():
    c = G["roi_inference"]
    boxes = O.apply_deltas(c["deltas"], c["props"], (10.0, 10.0, 5.0, 5.0))
    r, src = O.fast_rcnn_inference_single_image(boxes, torch.softmax(c["logits"], -1), c["image_shape"], 0.05, 0.5, 100,
                                                c["logits"], c["deltas"])
    assert torch.equal(src, c["src_idx"])
    assert torch.equal(r.pred_classes, c["pred_classes"])
    close(r.pred_boxes.tensor, c["pred_boxes"], 1e-5)
    close(r.scores, c["scores"])
    assert torch.equal(r.scores_logists, c["scores_logists"])
    assert torch.equal(r.boxes_sigma, c["boxes_sigma"])


def test_rpn_proposal_selection():
    c = G["rpn_proposals"]
    res = O.find_top_rpn_proposals(c["proposals"], c["logits"], c["image_sizes"], 0.7, 200, 50, 0.0, True,
                                   c["deltas"][..., 4:])
    for r, b, s in zip(res, c["boxes"], c["scores"]):
        close(r.proposal_boxes.tensor, b, 1e-5)
        close(r.objectness_logits, s)


def test_rpn_losses():
    c = G["rpn_losses"]
    for efl in (True, False):
        for tau in ((0.5, 0.5), (0.25, 0.25)):
            r = O.rpn_loss_unsupervised(c["logits"], c["soft"], c["deltas"], c["masks"], c["mgt"], c["sig"], c["anchors"],
                                        efl, (0.5, 0.5), tau, 256, (1.0, 1.0, 1.0, 1.0))
            ref = c[f"unsup_efl{int(efl)}_tau{tau[0]}"]
            close(r["loss_rpn_cls"], ref[0])
            close(r["loss_rpn_loc"], ref[1])
    r = O.rpn_losses(c["anchors"], c["logits"], c["labels"], c["deltas"], c["mgt"], 256, (1.0, 1.0, 1.0, 1.0))
    close(r["loss_rpn_cls"], c["sup"][0])
    close(r["loss_rpn_loc"], c["sup"][1])


class _Sampler:
    def __init__(self, prio):
        self.prio_ = prio

    def prio(self, tag, n):
        kind, i = tag
        return self.prio_[i][0 if kind.endswith("pos") else 1]


def test_anchor_labelling():
    c = G["rpn_labelling"]
    m = O.OracleRCNN.__new__(O.OracleRCNN)
    m.cfg = O.OracleCfg()
    m.sampler = _Sampler(c["prio"])
    insts = [O.OInst((192, 272), gt_boxes=O.OBoxes(b)) for b in c["gt_boxes"]]
    lab, mg = O.OracleRCNN.label_and_sample_anchors(m, c["anchors"], insts)
    for a, b in zip(lab, c["labels"]):
        assert torch.equal(a, b)
    for a, b, l in zip(mg, c["matched_gt"], lab):
        assert torch.equal(a[l == 1], b[l == 1])
    insts = [O.OInst((192, 272), pseudo_boxes=O.OBoxes(b), scores_logists=l, boxes_sigma=s)
             for b, l, s in zip(c["pseudo_boxes"], c["pseudo_logits"], c["pseudo_sigma"])]
    gl, am, mg2, ms = O.OracleRCNN.label_and_sample_anchors(m, c["anchors"], insts, True, True)
    for a, b in zip(gl, c["u_labels"]):
        assert torch.equal(a, b)
    for a, b in zip(am, c["u_masks"]):
        assert torch.equal(a, b)
    for a, b in zip(ms, c["u_sigma"]):
        assert torch.equal(a, b)
    for a, b, k in zip(mg2, c["u_matched"], am):
        assert torch.equal(a[k], b[k])


def test_differentiable_anchors():
    c = G["anchors"]
    a = O.grid_anchors(O.differentiable_cell_anchors(c["wh"]), c["H"], c["W"], 16, 0.0)
    assert torch.equal(a, c["anchors"])


def test_unsup_roi_sampling():
    c = G["roi_unsup_sampling"]
    l = G["rpn_labelling"]
    m = O.OracleRCNN.__new__(O.OracleRCNN)
    m.cfg = O.OracleCfg()
    props = [O.OInst((192, 272), proposal_boxes=O.OBoxes(p)) for p in c["props"]]
    tg = [O.OInst((192, 272), pseudo_boxes=O.OBoxes(b), scores_logists=lg, boxes_sigma=s)
          for b, lg, s in zip(l["pseudo_boxes"], l["pseudo_logits"], l["pseudo_sigma"])]
    res = O.OracleRCNN.label_and_sample_proposals(m, props, tg, "unsupervised")
    for r, b, p, s, g in zip(res, c["boxes"], c["pseudo"], c["soft"], c["sigma"]):
        assert torch.equal(r.proposal_boxes.tensor, b)
        assert torch.equal(r.pseudo_boxes.tensor, p)
        assert torch.equal(r.soft_label, s)
        assert torch.equal(r.boxes_sigma, g)


def test_vgg_block():
    c = G["vgg_block"]
    x = c["x"]
    for w, b in zip(c["w"], c["b"]):
        x = torch.relu(torch.nn.functional.conv2d(x, w, b, padding=1))
    x = torch.nn.functional.max_pool2d(x, 2, 2)
    close(x, c["y"], 1e-6)


def test_append_gt():
    c = G["append_gt"]
    assert torch.equal(torch.cat([c["props"], c["gt"]], 0), c["boxes"])
