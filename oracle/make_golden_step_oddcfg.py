"""Generates tests/golden/pt_reference_step_oddcfg_golden.pt: THREE consecutive post-burn-in steps of the REFERENCE'S OWN
`PTrainer.run_step` (pt/engine/trainer.py, imported unmodified; set-up of oracle/make_golden_step.py) with the
trainer-level hyper-parameters away from their defaults: SOURCE_LOSS_WEIGHT 0.5, TARGET_UNSUP_LOSS_WEIGHT 2.0,
EMA_KEEP_RATE 0.99, TEACHER_UPDATE_ITER 2 (so the teacher is copied at step 0, left alone at step 1 and EMA-updated at
step 2), SGD momentum 0.8, weight decay 5e-4, learning rate 0.004.

    python oracle/make_golden_step_oddcfg.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed."""
import os
import random
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_step as S  # noqa: E402

M, O, ref_trainer, d2shim_model = S.M, S.O, S.ref_trainer, S.d2shim_model
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

H, W, N, K, SEED = 128, 160, 2, 8, 51
NS = types.SimpleNamespace
TRAINER_CFG = dict(source_loss_weight=0.5, target_unsup_loss_weight=2.0, ema_keep_rate=0.99, teacher_update_iter=2,
                   momentum=0.8, weight_decay=5e-4, base_lr=0.004)


def main():
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.UNSUPNET.SOURCE_LOSS_WEIGHT = TRAINER_CFG["source_loss_weight"]
    cfg.UNSUPNET.TARGET_UNSUP_LOSS_WEIGHT = TRAINER_CFG["target_unsup_loss_weight"]
    cfg.UNSUPNET.EMA_KEEP_RATE = TRAINER_CFG["ema_keep_rate"]
    cfg.UNSUPNET.TEACHER_UPDATE_ITER = TRAINER_CFG["teacher_update_iter"]
    cfg.SOLVER.MOMENTUM, cfg.SOLVER.WEIGHT_DECAY, cfg.SOLVER.BASE_LR = \
        TRAINER_CFG["momentum"], TRAINER_CFG["weight_decay"], TRAINER_CFG["base_lr"]
    ocfg = O.OracleCfg(num_classes=K)
    sd = O.OracleRCNN(ocfg, seed=SEED).ref_state_dict()
    sd_t = O.OracleRCNN(ocfg, seed=SEED + 1).ref_state_dict()
    student = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    teacher = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd_t.items()})
    student.train()
    teacher.train()
    lab = O.synthetic_batch(N, H, W, K, 81, boxes_per_image=4)
    unl = O.synthetic_batch(N, H, W, K, 82, labelled=False)
    g = torch.Generator().manual_seed(79)
    R = (H // 16) * (W // 16) * 9
    L = cfg.MODEL.RPN.POST_NMS_TOPK_TRAIN + 16
    prio = {"rpn": (torch.rand(2 * N, R, generator=g), torch.rand(2 * N, R, generator=g)),
            "roi": (torch.rand(2 * N, L, generator=g), torch.rand(2 * N, L, generator=g))}
    d2shim_model.PRIO.provider = lambda tag, n: prio[tag[0].split("_")[0]][0 if tag[0].endswith("pos") else 1][tag[1]][:n]

    def batches():
        while True:
            yield (M.to_ref(lab), M.to_ref(lab), M.to_ref(unl), M.to_ref(unl))

    captured, draws = [], []
    real_uniform = random.uniform

    def recording_uniform(a, b):
        v = real_uniform(a, b)
        draws.append(v)
        return v
    random.seed(8)
    random.uniform = recording_uniform
    me = NS(cfg=cfg, model=student, model_teacher=teacher, iter=0,
            optimizer=torch.optim.SGD([p for p in student.parameters() if p.requires_grad], lr=cfg.SOLVER.BASE_LR,
                                      momentum=cfg.SOLVER.MOMENTUM, weight_decay=cfg.SOLVER.WEIGHT_DECAY),
            _trainer=NS(iter=0, _data_loader_iter=batches()))
    for name in ("resize", "_update_teacher_model", "process_pseudo_label", "threshold_bbox", "remove_label",
                 "add_label", "clip_gradient"):
        setattr(me, name, types.MethodType(getattr(ref_trainer.PTrainer, name), me))
    me._write_metrics = lambda md: captured.append({k: float(v) for k, v in md.items() if k.startswith("loss")})
    out = dict(H=H, W=W, N=N, K=K, seed=SEED, teacher_seed=SEED + 1, prio=prio, trainer_cfg=TRAINER_CFG,
               lab_images=[d["image"] for d in lab], unl_images=[d["image"] for d in unl],
               gt_boxes=[d["instances"].gt_boxes.tensor for d in lab],
               gt_classes=[d["instances"].gt_classes for d in lab], steps=[])
    for it in range(3):
        me.iter = it
        d2shim_model.PRIO.reset()
        n0 = len(draws)
        ref_trainer.PTrainer.run_step(me)
        out["steps"].append(dict(losses=captured[-1], ratios=list(draws[n0:]), student=S.sample_params(student),
                                 teacher=S.sample_params(teacher)))
        print("step", it, {k: round(v, 5) for k, v in captured[-1].items()})
    random.uniform = real_uniform
    t = [s["teacher"]["roi_heads.box_predictor.cls_score.weight"] for s in out["steps"]]
    assert torch.equal(t[0], t[1]) and not torch.equal(t[1], t[2])  # TEACHER_UPDATE_ITER = 2
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_step_oddcfg_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
