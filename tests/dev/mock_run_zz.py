"""Developer aid (CPU, not part of the suite): runs the model-level tests of tests/test_zz_next_rows_gpu.py against a
MOCK of the CUDA model that answers with the oracle's results wrapped in the package's containers. It cannot say
anything about the kernels; it checks the TEST CODE (keys, shapes, container handling, comparison logic) before the
first hardware run. Usage: python tests/dev/mock_run_zz.py"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pt_oracle as O  # noqa: E402
from probabilisticteacher_b200.modeling import postprocessing  # noqa: E402
from probabilisticteacher_b200.modeling.meta_arch import rcnn  # noqa: E402
from probabilisticteacher_b200.structures import Boxes, FreeInstances  # noqa: E402


class _Sampler:
    def __init__(self, model):
        self.model = model

    def prio(self, tag, n):
        grp, which = tag[0].split("_")
        pr = self.model.prio_override[grp]
        return (pr[0] if which == "pos" else pr[1])[tag[1]][:n].cpu()


class MockModel:
    """`build_model(cfg, device, precision=..., with_grads=...)` stand-in: oracle inside, package containers outside."""

    def __init__(self, cfg, device=None, **kw):
        m, u = cfg.MODEL, cfg.UNSUPNET
        r, h = m.RPN, m.ROI_HEADS
        self.ocfg = O.OracleCfg(
            num_classes=h.NUM_CLASSES, anchor_generator=m.ANCHOR_GENERATOR.NAME, pixel_std=tuple(m.PIXEL_STD),
            anchor_offset=m.ANCHOR_GENERATOR.OFFSET, rpn_nms_thresh=r.NMS_THRESH,
            rpn_pre_nms_topk=(r.PRE_NMS_TOPK_TRAIN, r.PRE_NMS_TOPK_TEST), rpn_post_nms_topk=(r.POST_NMS_TOPK_TRAIN, r.POST_NMS_TOPK_TEST),
            rpn_batch_per_image=r.BATCH_SIZE_PER_IMAGE, rpn_positive_fraction=r.POSITIVE_FRACTION,
            rpn_iou_thresholds=tuple(r.IOU_THRESHOLDS), roi_batch_per_image=h.BATCH_SIZE_PER_IMAGE,
            roi_positive_fraction=h.POSITIVE_FRACTION, roi_iou_threshold=h.IOU_THRESHOLDS[0],
            roi_score_thresh_test=h.SCORE_THRESH_TEST, roi_nms_thresh_test=h.NMS_THRESH_TEST,
            detections_per_image=cfg.TEST.DETECTIONS_PER_IMAGE, roi_bbox_weights=tuple(m.ROI_BOX_HEAD.BBOX_REG_WEIGHTS),
            efl=bool(u.EFL), tau=tuple(u.TAU), efl_lambda=tuple(u.EFL_LAMBDA))
        self.om = O.OracleRCNN(self.ocfg, seed=0)
        self.om.sampler = _Sampler(self)
        self.prio_override = None
        self.training = True

    def load_state_dict(self, sd, strict=True):
        self.om.load_ref_state_dict(sd)

    def init_synthetic(self, seed=0):
        self.om = O.OracleRCNN(self.ocfg, seed=seed)
        self.om.sampler = _Sampler(self)

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    @staticmethod
    def _to_oracle(batch):
        out = []
        for d in batch:
            nd = {"image": d["image"].cpu(), "height": d.get("height"), "width": d.get("width")}
            if "instances" in d:
                i = d["instances"]
                f = {k: (O.OBoxes(v.tensor.cpu()) if hasattr(v, "tensor") else v.cpu()) for k, v in i.get_fields().items()}
                nd["instances"] = O.OInst(tuple(i.image_size), **f)
            out.append(nd)
        return out

    @staticmethod
    def _free(o, fields):
        inst = FreeInstances(tuple(o.image_size))
        for k in fields:
            v = getattr(o, k)
            inst.set(k, Boxes(O._bt(v)) if k.endswith("boxes") else v)
        inst._count = torch.tensor(len(getattr(o, fields[-1])))
        return inst

    def inference(self, batch, do_postprocess=True):
        with torch.no_grad():
            _, _, roih, _ = self.om(self._to_oracle(batch), branch="unsup_data_weak", training=False)
        inst = [self._free(r, ("pred_boxes", "scores", "pred_classes", "scores_logists", "boxes_sigma")) for r in roih]
        if do_postprocess:
            return postprocessing.postprocess_batch(inst, batch, [tuple(i.image_size) for i in inst])
        return inst

    def __call__(self, batch, branch="supervised", danchor=False):
        if not self.training:
            return self.inference(batch)
        with torch.no_grad():
            losses, props, roih, _ = self.om(self._to_oracle(batch), branch=branch, danchor=danchor)
        if branch == "unsup_data_weak":
            return ({}, [self._free(p, ("proposal_boxes", "objectness_logits")) for p in props],
                    [self._free(r, ("pred_boxes", "scores", "pred_classes", "scores_logists", "boxes_sigma")) for r in roih], None)
        return losses, [], [], None


def main():
    rcnn.build_model = lambda cfg, device=None, **kw: MockModel(cfg, device, **kw)
    torch.cuda.synchronize = lambda *a, **k: None
    import test_zz_next_rows_gpu as Z
    cpu = torch.device("cpu")
    runs = [("eval golden c2f", lambda: Z.test_eval_mode_vs_reference_model_golden(cpu, "c2f_upscaled")),
            ("eval golden k1", lambda: Z.test_eval_mode_vs_reference_model_golden(cpu, "k1_default_anchors_mixed_sizes")),
            ("config1", lambda: Z.test_full_size_configs_vs_reference_model_golden(cpu, "config1")),
            ("config4", lambda: Z.test_full_size_configs_vs_reference_model_golden(cpu, "config4")),
            ("empty pseudo (one image)", lambda: Z.test_unsupervised_branch_without_pseudo_labels(cpu, "second_image_empty")),
            ("empty pseudo (all)", lambda: Z.test_unsupervised_branch_without_pseudo_labels(cpu, "all_empty")),
            ("no gt", lambda: Z.test_supervised_branch_without_any_ground_truth(cpu)),
            ("unsupnet variant 0", lambda: Z.test_unsupervised_branch_other_unsupnet_settings(cpu, 0)),
            ("unsupnet variant 1", lambda: Z.test_unsupervised_branch_other_unsupnet_settings(cpu, 1)),
            ("odd config", lambda: Z.test_every_hyper_parameter_away_from_its_default(cpu)),
            ("postprocess scaling", lambda: Z.test_eval_mode_postprocess_scaling(cpu))]
    failed = 0
    for name, fn in runs:
        try:
            fn()
            print("ok  ", name)
        except Exception as e:  # noqa: BLE001
            failed += 1
            import traceback
            print("FAIL", name, type(e).__name__, str(e)[:300])
            traceback.print_exc(limit=4)
    print("failed:", failed)


if __name__ == "__main__":
    main()
