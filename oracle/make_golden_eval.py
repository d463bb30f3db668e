"""Generates tests/golden/pt_reference_eval_golden.pt: EVAL-MODE inference by the REFERENCE'S OWN MODEL CLASSES --
`GuassianGeneralizedRCNN.forward` in eval mode (pt/modeling/meta_arch/rcnn.py:33-34) -> detectron2 v0.5
`GeneralizedRCNN.inference` (restated in oracle/d2shim_model.py) -> the reference's `GuassianRPN.forward` inference
branch with the TEST top-k (rpn.py:80-154, proposal_utils.py:27-154 with 6000 / 1000), `GuassianROIHead.forward`
eval branch (roi_heads.py:89-129,187-190) and `GuassianFastRCNNOutputLayers.inference` (fast_rcnn.py:338-409,34-120),
then `detector_postprocess` to an output size that differs from the network input size.

    python oracle/make_golden_eval.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed.
tests/test_oracle_golden_model.py replays it with oracle/pt_oracle.py + the package's `detector_postprocess`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_model as M  # noqa: E402

O = M.O
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

# (name, image sizes, requested output sizes, K, anchor generator, weight seed)
CASES = [("c2f_upscaled", [(128, 160), (128, 160)], [(256, 320), (192, 200)], 8, "DifferentiableAnchorGenerator", 17),
         ("k1_default_anchors_mixed_sizes", [(112, 160), (128, 144)], [(112, 160), (64, 72)], 1, "DefaultAnchorGenerator", 18)]


def run_case(sizes, out_sizes, K, anchor_name, seed):
    cfg = c2f_config()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = K
    cfg.MODEL.ANCHOR_GENERATOR.NAME = anchor_name
    ocfg = O.OracleCfg(num_classes=K, anchor_generator=anchor_name)
    sd = O.OracleRCNN(ocfg, seed=seed).ref_state_dict()
    model = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    model.eval()
    batch = []
    for i, ((h, w), (oh, ow)) in enumerate(zip(sizes, out_sizes)):
        d = O.synthetic_batch(1, h, w, K, 70 + i, labelled=False)[0]
        batch.append({"image": d["image"], "height": oh, "width": ow})
    with torch.no_grad():
        res = model(batch)
        raw = model.inference(batch, do_postprocess=False)
    dets = []
    for r, q in zip(res, raw):
        i = r["instances"]
        dets.append(dict(image_size=tuple(i.image_size), pred_boxes=i.pred_boxes.tensor.clone(), scores=i.scores.clone(),
                         pred_classes=i.pred_classes.clone(), scores_logists=i.scores_logists.clone(),
                         boxes_sigma=i.boxes_sigma.clone(), raw_boxes=q.pred_boxes.tensor.clone()))
        print("  image", tuple(q.image_size), "->", tuple(i.image_size), len(q.pred_boxes), "raw,", len(i.pred_boxes), "kept")
    return dict(sizes=sizes, out_sizes=out_sizes, K=K, anchor_generator=anchor_name, seed=seed,
                images=[b["image"] for b in batch], detections=dets)


def main():
    out = {}
    for name, sizes, out_sizes, K, ag, seed in CASES:
        print(name)
        out[name] = run_case(sizes, out_sizes, K, ag, seed)
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_eval_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
