"""Generates tests/golden/pt_reference_model_golden.pt: one post-burn-in iteration's forward passes computed by the
REFERENCE'S OWN MODEL CLASSES -- `GuassianGeneralizedRCNN.forward` (pt/modeling/meta_arch/rcnn.py:30-92) driving
the reference's `VGG` (built by `build_vgg_backbone`), `GuassianRPN` + `GuassianRPNHead`,
`DifferentiableAnchorGenerator`, `GuassianROIHead` and `GuassianFastRCNNOutputLayers`, all imported unmodified from
/root/reference and constructed on top of the functional detectron2 v0.5 base classes of oracle/d2shim_model.py.
Run here (the reference tree does not exist on the GPU box):

    python oracle/make_golden_model.py

Stored: the inputs (uint8 images, ground truth, sampling priorities, the seed of the weights) and the outputs of
the three branches (supervised losses; teacher RPN proposals and pseudo labels; unsupervised losses and the
gradient that reaches the differentiable anchors). tests/test_oracle_golden_model.py runs oracle/pt_oracle.py's
`OracleRCNN` on the same inputs and weights and compares: this pins the ORCHESTRATION of the oracle (what is
called in which order with which arguments) to the reference's own forward methods; the function-level fixture
(oracle/make_golden.py) pins the arithmetic.
"""
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import d2shim_model  # noqa: E402

B = d2shim_model.install()
torch.Tensor.cuda = lambda self, *a, **k: self  # DifferentiableAnchorGenerator.__init__ calls .cuda() (anchor_generator.py:69)

from detectron2.modeling.matcher import Matcher  # noqa: E402
from detectron2.structures import Boxes  # noqa: E402
from pt.modeling import anchor_generator as ref_ag  # noqa: E402
from pt.modeling.backbone import vgg as ref_vgg  # noqa: E402
from pt.modeling.box_regression import Box2BoxTransform  # noqa: E402
from pt.modeling.meta_arch import rcnn as ref_rcnn  # noqa: E402
from pt.modeling.proposal_generator import rpn as ref_rpn  # noqa: E402
from pt.modeling.roi_heads import fast_rcnn as ref_fr  # noqa: E402
from pt.modeling.roi_heads import roi_heads as ref_rh  # noqa: E402
from pt.structures.instances import FreeInstances  # noqa: E402

from oracle import pt_oracle as O  # noqa: E402
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

# (name, image sizes, K, anchor generator, weight seed, ground-truth boxes per image)
CASES = [("c2f", [(128, 160), (128, 160)], 8, "DifferentiableAnchorGenerator", 11, (4, 4)),
         ("k1_default_anchors_mixed_sizes", [(112, 160), (128, 144)], 1, "DefaultAnchorGenerator", 12, (4, 4)),
         ("c2f_image_without_gt", [(128, 160), (128, 160)], 8, "DifferentiableAnchorGenerator", 13, (3, 0))]


def build_reference_model(cfg, sd):
    """The reference's classes, constructed with the arguments their `from_config` methods would pass
    (d2 v0.5 from_config + rpn.py:71-78, roi_heads.py:46-87, fast_rcnn.py:171-177, vgg.py:189-230)."""
    with tempfile.TemporaryDirectory() as tmp:
        # vgg.py:127-152 loads an ImageNet checkpoint unconditionally: give it one with torchvision's key names
        names = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
        convs = [k[:-7] for k in sd if k.startswith("backbone.") and k.endswith(".weight")]
        ck = {}
        for n_, c in zip(names, convs):
            ck[f"features.{n_}.weight"] = sd[c + ".weight"]
            ck[f"features.{n_}.bias"] = sd[c + ".bias"]
        path = os.path.join(tmp, "vgg16_caffe.pth")
        torch.save(ck, path)
        cfg.MODEL.VGG.PRETRAIN = path
        backbone = ref_vgg.build_vgg_backbone(cfg, B["ShapeSpec"](channels=3))
    a = cfg.MODEL.ANCHOR_GENERATOR
    if a.NAME == "DifferentiableAnchorGenerator":
        anchor_gen = ref_ag.DifferentiableAnchorGenerator(anchor=a.ANCHOR, strides=[16], offset=a.OFFSET)
    else:  # configs/Guassian-RCNN-VGG.yaml:9-12 (detectron2's own generator, restated in d2shim_model)
        anchor_gen = B["DefaultAnchorGenerator"](sizes=a.SIZES, aspect_ratios=a.ASPECT_RATIOS, strides=[16],
                                                 offset=a.OFFSET)
    head = ref_rpn.GuassianRPNHead(in_channels=512, num_anchors=9, box_dim=8)  # box_dim doubled, rpn.py:50-55
    r = cfg.MODEL.RPN
    rpn = ref_rpn.GuassianRPN(
        cfg=cfg, in_features=["vgg_block5"], head=head, anchor_generator=anchor_gen,
        anchor_matcher=Matcher(list(r.IOU_THRESHOLDS), list(r.IOU_LABELS), allow_low_quality_matches=True),
        box2box_transform=Box2BoxTransform(weights=tuple(r.BBOX_REG_WEIGHTS)),
        batch_size_per_image=r.BATCH_SIZE_PER_IMAGE, positive_fraction=r.POSITIVE_FRACTION,
        pre_nms_topk=(r.PRE_NMS_TOPK_TRAIN, r.PRE_NMS_TOPK_TEST), post_nms_topk=(r.POST_NMS_TOPK_TRAIN, r.POST_NMS_TOPK_TEST),
        nms_thresh=r.NMS_THRESH, min_box_size=cfg.MODEL.PROPOSAL_GENERATOR.MIN_SIZE, anchor_boundary_thresh=-1,
        loss_weight={"loss_rpn_cls": r.LOSS_WEIGHT, "loss_rpn_loc": r.LOSS_WEIGHT},
        box_reg_loss_type="smooth_l1", smooth_l1_beta=0.0)
    h = cfg.MODEL.ROI_HEADS
    box_head = B["FastRCNNConvFCHead"](B["ShapeSpec"](channels=512, height=7, width=7), cfg.MODEL.ROI_BOX_HEAD.FC_DIM, 2)
    predictor = ref_fr.GuassianFastRCNNOutputLayers(
        cfg=cfg, model_type=cfg.UNSUPNET.MODEL_TYPE, input_shape=box_head.output_shape,
        box2box_transform=Box2BoxTransform(weights=tuple(cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS)),
        num_classes=h.NUM_CLASSES, test_score_thresh=h.SCORE_THRESH_TEST, test_nms_thresh=h.NMS_THRESH_TEST,
        test_topk_per_image=cfg.TEST.DETECTIONS_PER_IMAGE, cls_agnostic_bbox_reg=False, smooth_l1_beta=0.0,
        box_reg_loss_type="smooth_l1", loss_weight=1.0)
    roi_heads = ref_rh.GuassianROIHead(
        cfg=cfg, box_in_features=["vgg_block5"], box_pooler=B["ROIPooler"](7, (1.0 / 16,), 0, "ROIAlignV2"),
        box_head=box_head, box_predictor=predictor, num_classes=h.NUM_CLASSES,
        batch_size_per_image=h.BATCH_SIZE_PER_IMAGE, positive_fraction=h.POSITIVE_FRACTION,
        proposal_matcher=Matcher(list(h.IOU_THRESHOLDS), list(h.IOU_LABELS), allow_low_quality_matches=False),
        proposal_append_gt=h.PROPOSAL_APPEND_GT, train_on_pred_boxes=False)
    model = ref_rcnn.GuassianGeneralizedRCNN(backbone=backbone, proposal_generator=rpn, roi_heads=roi_heads,
                                             pixel_mean=cfg.MODEL.PIXEL_MEAN, pixel_std=cfg.MODEL.PIXEL_STD)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    return model


def to_ref(batch):
    out = []
    for d in batch:
        nd = {"image": d["image"], "height": d["height"], "width": d["width"]}
        if "instances" in d:
            i = d["instances"]
            nd["instances"] = FreeInstances(i.image_size, gt_boxes=Boxes(i.gt_boxes.tensor.clone()),
                                            gt_classes=i.gt_classes.clone())
        out.append(nd)
    return out


def run_case(sizes, K, anchor_name, seed, n_gt):
    cfg = c2f_config()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = K
    cfg.MODEL.ANCHOR_GENERATOR.NAME = anchor_name
    ocfg = O.OracleCfg(num_classes=K, anchor_generator=anchor_name)
    sd = O.OracleRCNN(ocfg, seed=seed).ref_state_dict()  # weights are inputs: seeded initialisers, not stored
    model = build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    model.train()
    N = len(sizes)
    lab = [O.synthetic_batch(1, h, w, K, 5 + i, boxes_per_image=n_gt[i])[0] for i, (h, w) in enumerate(sizes)]
    unl = [O.synthetic_batch(1, h, w, K, 50 + i, labelled=False)[0] for i, (h, w) in enumerate(sizes)]
    g = torch.Generator().manual_seed(99)
    Hm, Wm = max(s[0] for s in sizes), max(s[1] for s in sizes)
    R = (Hm // 16) * (Wm // 16) * 9
    L = cfg.MODEL.RPN.POST_NMS_TOPK_TRAIN + 16
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}

    def provider(tag, n):
        grp, which = tag[0].split("_")
        return prio[grp][0 if which == "pos" else 1][tag[1]][:n]
    d2shim_model.PRIO.provider = provider

    out = dict(sizes=sizes, N=N, K=K, seed=seed, anchor_generator=anchor_name, prio=prio,
               lab_images=[d["image"] for d in lab], unl_images=[d["image"] for d in unl],
               gt_boxes=[d["instances"].gt_boxes.tensor for d in lab],
               gt_classes=[d["instances"].gt_classes for d in lab])

    with torch.no_grad():
        d2shim_model.PRIO.reset()
        losses, _, _, _ = model(to_ref(lab), branch="supervised")
        out["sup_losses"] = {k: v.clone() for k, v in losses.items()}
        d2shim_model.PRIO.reset()
        _, prop_rpn, prop_roih, roi_pred = model(to_ref(unl), branch="unsup_data_weak")
    out["teacher_rpn_boxes"] = [p.proposal_boxes.tensor.clone() for p in prop_rpn]
    out["teacher_rpn_logits"] = [p.objectness_logits.clone() for p in prop_rpn]
    out["teacher_roih"] = [dict(pred_boxes=p.pred_boxes.tensor.clone(), scores=p.scores.clone(),
                                pred_classes=p.pred_classes.clone(), scores_logists=p.scores_logists.clone(),
                                boxes_sigma=p.boxes_sigma.clone()) for p in prop_roih]
    # pt/engine/trainer.py:179-257: the teacher's detections become the pseudo labels of the strong view
    unl_q = []
    for d, p in zip(to_ref(unl), prop_roih):
        inst = FreeInstances(p.image_size, pseudo_boxes=Boxes(p.pred_boxes.tensor.clone()),
                             scores_logists=p.scores_logists.clone(), boxes_sigma=p.boxes_sigma.clone())
        unl_q.append(dict(d, instances=inst))
    differentiable = anchor_name == "DifferentiableAnchorGenerator"
    for danchor, key in ((True, "unsup_anchor_grad"), (False, "unsup_anchor_grad_no_danchor")):
        model.zero_grad()
        d2shim_model.PRIO.reset()
        losses, _, _, _ = model(unl_q, branch="unsupervised", danchor=danchor)
        if danchor:
            out["unsup_losses"] = {k: v.detach().clone() for k, v in losses.items()}
        sum(losses.values()).backward()
        if differentiable:
            ag = model.proposal_generator.anchor_generator.anchor_0.grad
            out[key] = torch.zeros(9, 2) if ag is None else ag.clone()
    # one trainable tensor's gradient of the supervised branch (autograd through the reference's own graph)
    model.zero_grad()
    d2shim_model.PRIO.reset()
    losses, _, _, _ = model(to_ref(lab), branch="supervised")
    sum(losses.values()).backward()
    out["sup_grad_rpn_objectness_w"] = model.proposal_generator.rpn_head.objectness_logits.weight.grad.clone()
    out["sup_grad_cls_score_w"] = model.roi_heads.box_predictor.cls_score.weight.grad.clone()
    out["sup_grad_conv5_3_b"] = model.backbone.vgg_block5[0].conv3.bias.grad.clone()

    for k, v in out["sup_losses"].items():
        print("  sup", k, float(v))
    for k, v in out["unsup_losses"].items():
        print("  unsup", k, float(v))
    print("  teacher proposals", [len(b) for b in out["teacher_rpn_boxes"]], "detections",
          [len(r["scores"]) for r in out["teacher_roih"]])
    return out


def main():
    out = {}
    for name, sizes, K, anchor_name, seed, n_gt in CASES:
        print(name)
        out[name] = run_case(sizes, K, anchor_name, seed, n_gt)
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_model_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
