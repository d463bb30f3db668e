"""Generates tests/golden/pt_reference_step_golden.pt: two consecutive post-burn-in TRAINING STEPS executed by the
REFERENCE'S OWN `PTrainer.run_step` (pt/engine/trainer.py:263-392, imported unmodified) on a namespace object that
carries the reference's own model classes (oracle/make_golden_model.py) for student and teacher, a torch SGD
optimizer built as detectron2's `build_optimizer` does, and the reference's own `resize`,
`_update_teacher_model`, `process_pseudo_label`, `threshold_bbox`, `remove_label`, `add_label` and
`clip_gradient` methods. Step 0 is the first post-burn-in iteration (teacher <- copy of the student,
trainer.py:293-295), step 1 an EMA iteration (keep rate 0.9996).

    python oracle/make_golden_step.py

Stored: inputs (images, ground truth, sampling priorities, weight seed, the `random.uniform` draws of `resize`),
the 8 losses of each step, and 64 fixed samples of every student / teacher parameter after each step.
tests/test_oracle_golden_step.py replays the two steps with oracle/pt_oracle.py (`run_step`) and compares."""
import copy
import os
import random
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_model as M  # noqa: E402  (installs the functional d2 base classes, imports pt.modeling)
from oracle import d2shim_any, d2shim_model  # noqa: E402

d2shim_any.install()
import pt.engine.trainer as ref_trainer  # noqa: E402

from oracle import pt_oracle as O  # noqa: E402
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

H, W, N, K, SEED = 128, 160, 2, 8, 21
NS = types.SimpleNamespace


def sample_idx(numel, n=64):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def sample_params(model):
    return {k: v.detach().reshape(-1)[sample_idx(v.numel())].clone() for k, v in model.state_dict().items()}


def main():
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    ocfg = O.OracleCfg(num_classes=K)
    sd = O.OracleRCNN(ocfg, seed=SEED).ref_state_dict()
    sd_t = O.OracleRCNN(ocfg, seed=SEED + 1).ref_state_dict()  # teacher starts different: step 0 must overwrite it
    student = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    teacher = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd_t.items()})
    student.train()
    teacher.train()

    lab = O.synthetic_batch(N, H, W, K, 31, boxes_per_image=4)
    unl = O.synthetic_batch(N, H, W, K, 32, labelled=False)
    g = torch.Generator().manual_seed(77)
    R = (H // 16) * (W // 16) * 9
    L = cfg.MODEL.RPN.POST_NMS_TOPK_TRAIN + 16
    prio = {"rpn": (torch.rand(2 * N, R, generator=g), torch.rand(2 * N, R, generator=g)),
            "roi": (torch.rand(2 * N, L, generator=g), torch.rand(2 * N, L, generator=g))}

    def provider(tag, n):
        grp, which = tag[0].split("_")
        return prio[grp][0 if which == "pos" else 1][tag[1]][:n]
    d2shim_model.PRIO.provider = provider

    def batches():
        while True:  # (label_q, label_k, unlabel_q, unlabel_k): strong / weak views are the same images here
            yield (M.to_ref(lab), M.to_ref(lab), M.to_ref(unl), M.to_ref(unl))

    captured = []
    draws = []
    real_uniform = random.uniform

    def recording_uniform(a, b):
        v = real_uniform(a, b)
        draws.append(v)
        return v
    random.seed(5)
    random.uniform = recording_uniform

    me = NS(cfg=cfg, model=student, model_teacher=teacher, iter=0,
            optimizer=torch.optim.SGD([p for p in student.parameters() if p.requires_grad], lr=cfg.SOLVER.BASE_LR,
                                      momentum=cfg.SOLVER.MOMENTUM, weight_decay=cfg.SOLVER.WEIGHT_DECAY),
            _trainer=NS(iter=0, _data_loader_iter=batches()))
    for name in ("resize", "_update_teacher_model", "process_pseudo_label", "threshold_bbox", "remove_label",
                 "add_label", "clip_gradient"):
        setattr(me, name, types.MethodType(getattr(ref_trainer.PTrainer, name), me))
    me._write_metrics = lambda md: captured.append({k: float(v) for k, v in md.items() if k.startswith("loss")})

    out = dict(H=H, W=W, N=N, K=K, seed=SEED, teacher_seed=SEED + 1, prio=prio, lr=cfg.SOLVER.BASE_LR,
               lab_images=[d["image"] for d in lab], unl_images=[d["image"] for d in unl],
               gt_boxes=[d["instances"].gt_boxes.tensor for d in lab],
               gt_classes=[d["instances"].gt_classes for d in lab], steps=[])
    for it in range(2):
        me.iter = it
        d2shim_model.PRIO.reset()
        n0 = len(draws)
        ref_trainer.PTrainer.run_step(me)
        out["steps"].append(dict(losses=captured[-1], ratios=list(draws[n0:]), student=sample_params(student),
                                 teacher=sample_params(teacher)))
        print("step", it, {k: round(v, 5) for k, v in captured[-1].items()}, "ratios", [round(r, 4) for r in draws[n0:]])
    random.uniform = real_uniform
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_step_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
