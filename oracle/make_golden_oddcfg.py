"""Generates tests/golden/pt_reference_oddcfg_golden.pt: one post-burn-in iteration's forward passes of the REFERENCE'S
OWN MODEL CLASSES under a configuration in which EVERY honoured hyper-parameter differs from the defaults (pixel std,
anchor offset, RPN / ROI thresholds, batch sizes, fractions, top-k sizes, score / NMS thresholds of the pseudo-label
filter, detections per image, box-regression weights): a path that ignored one of them cannot reproduce these numbers.

    python oracle/make_golden_oddcfg.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_model as M  # noqa: E402

O, d2shim_model = M.O, M.d2shim_model
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

H, W, K, N, WEIGHT_SEED, LAB_SEED, UNL_SEED, PRIO_SEED = 128, 160, 8, 2, 31, 71, 72, 98
# (config key, value) pairs applied with cfg.merge_from_list-like assignment; tests apply the same list
OVERRIDES = [("MODEL.PIXEL_STD", [57.375, 57.12, 58.395]), ("MODEL.ANCHOR_GENERATOR.OFFSET", 0.5),
             ("MODEL.RPN.NMS_THRESH", 0.6), ("MODEL.RPN.PRE_NMS_TOPK_TRAIN", 300), ("MODEL.RPN.POST_NMS_TOPK_TRAIN", 50),
             ("MODEL.RPN.BATCH_SIZE_PER_IMAGE", 64), ("MODEL.RPN.POSITIVE_FRACTION", 0.5),
             ("MODEL.RPN.IOU_THRESHOLDS", [0.2, 0.6]), ("MODEL.ROI_HEADS.BATCH_SIZE_PER_IMAGE", 128),
             ("MODEL.ROI_HEADS.POSITIVE_FRACTION", 0.5), ("MODEL.ROI_HEADS.IOU_THRESHOLDS", [0.4]),
             ("MODEL.ROI_HEADS.SCORE_THRESH_TEST", 0.1), ("MODEL.ROI_HEADS.NMS_THRESH_TEST", 0.4),
             ("TEST.DETECTIONS_PER_IMAGE", 20), ("MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS", (5.0, 5.0, 2.0, 2.0))]
ORACLE_KW = dict(pixel_std=(57.375, 57.12, 58.395), anchor_offset=0.5, rpn_nms_thresh=0.6, rpn_pre_nms_topk=(300, 6000),
                 rpn_post_nms_topk=(50, 1000), rpn_batch_per_image=64, rpn_positive_fraction=0.5,
                 rpn_iou_thresholds=(0.2, 0.6), roi_batch_per_image=128, roi_positive_fraction=0.5, roi_iou_threshold=0.4,
                 roi_score_thresh_test=0.1, roi_nms_thresh_test=0.4, detections_per_image=20,
                 roi_bbox_weights=(5.0, 5.0, 2.0, 2.0))


def apply_overrides(cfg, overrides):
    for key, value in overrides:
        node = cfg
        *parents, leaf = key.split(".")
        for p in parents:
            node = node[p]
        node[leaf] = value
    return cfg


def main():
    cfg = apply_overrides(c2f_config(), OVERRIDES)
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K, **ORACLE_KW), seed=WEIGHT_SEED).ref_state_dict()
    model = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    model.train()
    lab = O.synthetic_batch(N, H, W, K, LAB_SEED, boxes_per_image=4)
    unl = O.synthetic_batch(N, H, W, K, UNL_SEED, labelled=False)
    g = torch.Generator().manual_seed(PRIO_SEED)
    R, L = (H // 16) * (W // 16) * 9, 50 + 16
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}
    d2shim_model.PRIO.provider = lambda tag, n: prio[tag[0].split("_")[0]][0 if tag[0].endswith("pos") else 1][tag[1]][:n]
    out = dict(H=H, W=W, K=K, N=N, weight_seed=WEIGHT_SEED, lab_seed=LAB_SEED, unl_seed=UNL_SEED, prio_seed=PRIO_SEED,
               roi_prio_len=L, overrides=OVERRIDES, oracle_kw=ORACLE_KW)
    with torch.no_grad():
        d2shim_model.PRIO.reset()
        losses, _, _, _ = model(M.to_ref(lab), branch="supervised")
        out["sup_losses"] = {k: float(v) for k, v in losses.items()}
        d2shim_model.PRIO.reset()
        _, prop_rpn, prop_roih, _ = model(M.to_ref(unl), branch="unsup_data_weak")
        out["teacher_rpn_boxes"] = [p.proposal_boxes.tensor.clone() for p in prop_rpn]
        out["teacher_rpn_logits"] = [p.objectness_logits.clone() for p in prop_rpn]
        out["teacher_roih"] = [dict(pred_boxes=p.pred_boxes.tensor.clone(), scores=p.scores.clone(),
                                    pred_classes=p.pred_classes.clone(), scores_logists=p.scores_logists.clone(),
                                    boxes_sigma=p.boxes_sigma.clone()) for p in prop_roih]
        q = [dict(d, instances=M.FreeInstances(p.image_size, pseudo_boxes=M.Boxes(p.pred_boxes.tensor.clone()),
                                               scores_logists=p.scores_logists.clone(), boxes_sigma=p.boxes_sigma.clone()))
             for d, p in zip(M.to_ref(unl), prop_roih)]
        d2shim_model.PRIO.reset()
        losses, _, _, _ = model(q, branch="unsupervised", danchor=True)
        out["unsup_losses"] = {k: float(v) for k, v in losses.items()}
    print("sup", out["sup_losses"])
    print("teacher proposals", [len(b) for b in out["teacher_rpn_boxes"]], "detections", [len(r["scores"]) for r in out["teacher_roih"]])
    print("unsup", out["unsup_losses"])
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_oddcfg_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
