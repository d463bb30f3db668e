"""Generates tests/golden/pt_reference_burnin_golden.pt: two SOURCE-ONLY (burn-in, iter < BURN_UP_STEP) training
steps executed by the REFERENCE'S OWN `PTrainer.run_step` (pt/engine/trainer.py:263-290,379-386, imported
unmodified) on the reference's own model classes, in the set-up of oracle/make_golden_step.py; plus the parameter
name each `features.N.*` entry of a torchvision-style VGG16 file ends up under when the reference's OWN `VGG`
constructor loads it (pt/modeling/backbone/vgg.py:127-152) -- the table `checkpoint.vgg16_caffe_key_map` restates.

    python oracle/make_golden_burnin.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed.
tests/test_oracle_golden_step.py replays the steps with oracle/pt_oracle.py (`run_step_burn_in`)."""
import os
import random
import sys
import tempfile
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_step as S  # noqa: E402  (installs the shims, imports pt.engine.trainer unmodified)

M, O, ref_trainer, d2shim_model = S.M, S.O, S.ref_trainer, S.d2shim_model
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

H, W, N, K, SEED = 128, 160, 2, 8, 41
NS = types.SimpleNamespace


def vgg_file_mapping():
    """Builds the reference's VGG from a file whose 26 tensors are all different and reports, for every backbone
    parameter, which file entry it now equals."""
    g = torch.Generator().manual_seed(3)
    chans = [3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512]
    idx = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
    ck = {}
    for i, n_ in enumerate(idx):
        ck[f"features.{n_}.weight"] = torch.randn(chans[i + 1], chans[i], 3, 3, generator=g)
        ck[f"features.{n_}.bias"] = torch.randn(chans[i + 1], generator=g)
    cfg = c2f_config()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "vgg16_caffe.pth")
        torch.save(ck, path)
        cfg.MODEL.VGG.PRETRAIN = path
        backbone = M.ref_vgg.build_vgg_backbone(cfg, M.B["ShapeSpec"](channels=3))
    out = {}
    for name, p in backbone.state_dict().items():
        hits = [k for k, v in ck.items() if v.shape == p.shape and torch.equal(v, p)]
        assert len(hits) == 1, (name, hits)
        out[hits[0]] = name
    assert len(out) == 26
    return out


def main():
    cfg = c2f_config()
    assert cfg.UNSUPNET.BURN_UP_STEP == 4000  # iterations 0 and 1 are source-only
    ocfg = O.OracleCfg(num_classes=K)
    sd = O.OracleRCNN(ocfg, seed=SEED).ref_state_dict()
    sd_t = O.OracleRCNN(ocfg, seed=SEED + 1).ref_state_dict()
    student = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    teacher = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd_t.items()})
    student.train()
    teacher.train()
    teacher_before = S.sample_params(teacher)

    lab_q = O.synthetic_batch(N, H, W, K, 61, boxes_per_image=4)   # strong view
    lab_k = O.synthetic_batch(N, H, W, K, 62, boxes_per_image=3)   # weak view: different images, so order matters
    unl = O.synthetic_batch(N, H, W, K, 63, labelled=False)
    g = torch.Generator().manual_seed(78)
    R = (H // 16) * (W // 16) * 9
    L = cfg.MODEL.RPN.POST_NMS_TOPK_TRAIN + 16
    prio = {"rpn": (torch.rand(2 * N, R, generator=g), torch.rand(2 * N, R, generator=g)),
            "roi": (torch.rand(2 * N, L, generator=g), torch.rand(2 * N, L, generator=g))}

    def provider(tag, n):
        grp, which = tag[0].split("_")
        return prio[grp][0 if which == "pos" else 1][tag[1]][:n]
    d2shim_model.PRIO.provider = provider

    def batches():
        while True:
            yield (M.to_ref(lab_q), M.to_ref(lab_k), M.to_ref(unl), M.to_ref(unl))

    captured, draws = [], []
    real_uniform = random.uniform

    def recording_uniform(a, b):
        v = real_uniform(a, b)
        draws.append(v)
        return v
    random.seed(6)
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
               lab_q_images=[d["image"] for d in lab_q], lab_k_images=[d["image"] for d in lab_k],
               gt_boxes_q=[d["instances"].gt_boxes.tensor for d in lab_q],
               gt_classes_q=[d["instances"].gt_classes for d in lab_q],
               gt_boxes_k=[d["instances"].gt_boxes.tensor for d in lab_k],
               gt_classes_k=[d["instances"].gt_classes for d in lab_k], steps=[],
               vgg16_caffe_mapping=vgg_file_mapping())
    for it in range(2):
        me.iter = it
        d2shim_model.PRIO.reset()
        n0 = len(draws)
        ref_trainer.PTrainer.run_step(me)
        out["steps"].append(dict(losses=captured[-1], ratios=list(draws[n0:]), student=S.sample_params(student)))
        print("step", it, {k: round(v, 5) for k, v in captured[-1].items()}, "ratios", [round(r, 4) for r in draws[n0:]])
    random.uniform = real_uniform
    teacher_after = S.sample_params(teacher)
    assert all(torch.equal(teacher_before[k], teacher_after[k]) for k in teacher_before)  # burn-in never touches it
    assert sorted(captured[-1]) == ["loss_box_reg", "loss_cls", "loss_rpn_cls", "loss_rpn_loc"]
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_burnin_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
