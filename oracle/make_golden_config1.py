"""Generates tests/golden/pt_reference_config1_golden.pt: BASELINE.json config 1 at FULL SIZE -- the forward passes of
one post-burn-in iteration (supervised, teacher, unsupervised with `danchor`) on 1 source + 1 target synthetic
3x800x1333 image with `configs/Guassian-RCNN-VGG.yaml`'s model (DefaultAnchorGenerator, K = 8), computed by the
REFERENCE'S OWN MODEL CLASSES (set-up of oracle/make_golden_model.py, CPU fp32).

    python oracle/make_golden_config1.py        # ~1.5 min on 8 cores

Only seeds and results are stored (the two 3.2 MB images are regenerated from their seeds by the tests):
the 4 + 4 loss scalars, the teacher's post-NMS RPN proposals and its 100 pseudo labels.
Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_model as M  # noqa: E402

O, d2shim_model = M.O, M.d2shim_model
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

# python oracle/make_golden_config1.py config4  -> tests/golden/pt_reference_config4_golden.pt: BASELINE config 4's
# model and size (configs/pt/final_k2c.yaml: K = 1, train.sh's DifferentiableAnchorGenerator, one 3x600x2000 pair)
CASE = sys.argv[1] if len(sys.argv) > 1 else "config1"
if CASE == "config1":
    H, W, K, N = 800, 1333, 8, 1
    WEIGHT_SEED, LAB_SEED, UNL_SEED, PRIO_SEED = 23, 1, 2, 7
    ANCHORS = "DefaultAnchorGenerator"
elif CASE == "config4":
    H, W, K, N = 600, 2000, 1, 1
    WEIGHT_SEED, LAB_SEED, UNL_SEED, PRIO_SEED = 29, 3, 4, 9
    ANCHORS = "DifferentiableAnchorGenerator"
else:
    raise SystemExit(f"unknown case {CASE!r}")


def prios():
    """Same construction as tests/test_parity_x3_gpu.py::_prios."""
    g = torch.Generator().manual_seed(PRIO_SEED)
    R = (H // 16) * (W // 16) * 9
    L = 2000 + 16
    return {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}


def main():
    cfg = c2f_config()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = K
    cfg.MODEL.ANCHOR_GENERATOR.NAME = ANCHORS
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K, anchor_generator=ANCHORS), seed=WEIGHT_SEED).ref_state_dict()
    model = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    model.train()
    lab = O.synthetic_batch(N, H, W, K, LAB_SEED)
    unl = O.synthetic_batch(N, H, W, K, UNL_SEED, labelled=False)
    prio = prios()

    def provider(tag, n):
        grp, which = tag[0].split("_")
        return prio[grp][0 if which == "pos" else 1][tag[1]][:n]
    d2shim_model.PRIO.provider = provider

    out = dict(H=H, W=W, K=K, N=N, weight_seed=WEIGHT_SEED, lab_seed=LAB_SEED, unl_seed=UNL_SEED, prio_seed=PRIO_SEED,
               anchor_generator=ANCHORS)
    with torch.no_grad():
        d2shim_model.PRIO.reset()
        losses, _, _, _ = model(M.to_ref(lab), branch="supervised")
        out["sup_losses"] = {k: float(v) for k, v in losses.items()}
        print("sup", out["sup_losses"])
        d2shim_model.PRIO.reset()
        _, prop_rpn, prop_roih, _ = model(M.to_ref(unl), branch="unsup_data_weak")
        out["teacher_rpn_boxes"] = [p.proposal_boxes.tensor.clone() for p in prop_rpn]
        out["teacher_rpn_logits"] = [p.objectness_logits.clone() for p in prop_rpn]
        out["teacher_roih"] = [dict(pred_boxes=p.pred_boxes.tensor.clone(), scores=p.scores.clone(),
                                    pred_classes=p.pred_classes.clone(), scores_logists=p.scores_logists.clone(),
                                    boxes_sigma=p.boxes_sigma.clone()) for p in prop_roih]
        print("teacher proposals", [len(b) for b in out["teacher_rpn_boxes"]], "detections",
              [len(r["scores"]) for r in out["teacher_roih"]])
        unl_q = []
        for d, p in zip(M.to_ref(unl), prop_roih):  # pt/engine/trainer.py:179-257
            inst = M.FreeInstances(p.image_size, pseudo_boxes=M.Boxes(p.pred_boxes.tensor.clone()),
                                   scores_logists=p.scores_logists.clone(), boxes_sigma=p.boxes_sigma.clone())
            unl_q.append(dict(d, instances=inst))
        d2shim_model.PRIO.reset()
        losses, _, _, _ = model(unl_q, branch="unsupervised", danchor=True)
        out["unsup_losses"] = {k: float(v) for k, v in losses.items()}
        print("unsup", out["unsup_losses"])
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), f"pt_reference_{CASE}_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
