"""CPU: two whole training steps of the oracle against the REFERENCE'S OWN `PTrainer.run_step`.

tests/golden/pt_reference_step_golden.pt was produced by oracle/make_golden_step.py: pt/engine/trainer.py imported
unmodified, `PTrainer.run_step` executed twice (iteration == BURN_UP_STEP: teacher <- student copy; then an EMA
iteration) with the reference's own model classes, `resize`, pseudo-label repack, `clip_gradient` and EMA, and a
torch SGD optimizer. The oracle (`oracle.pt_oracle.run_step`) replays both steps from the same weights, images,
`resize` ratios and sampling priorities and must reproduce the 8 losses and the sampled student / teacher parameters
after each step."""
import os

import torch

from oracle import pt_oracle as O

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_step_golden.pt"), weights_only=False)


class _Sampler:
    def __init__(self, pr):
        self.pr = pr

    def prio(self, tag, n):
        grp, which = tag[0].split("_")
        return self.pr[grp][0 if which == "pos" else 1][tag[1]][:n]


def _sample_idx(numel, n=64):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def _batch():
    H, W = G["H"], G["W"]
    lab = [{"image": im.clone(), "height": H, "width": W,
            "instances": O.OInst((H, W), gt_boxes=O.OBoxes(b.clone()), gt_classes=c.clone())}
           for im, b, c in zip(G["lab_images"], G["gt_boxes"], G["gt_classes"])]
    unl = [{"image": im.clone(), "height": H, "width": W} for im in G["unl_images"]]
    return lab, unl


def test_two_training_steps_match_the_reference_trainer():
    cfg = O.OracleCfg(num_classes=G["K"], base_lr=G["lr"])
    student = O.OracleRCNN(cfg, seed=G["seed"])
    teacher = O.OracleRCNN(cfg, seed=G["teacher_seed"])
    student.sampler = _Sampler(G["prio"])
    opt = O.make_optimizer(student, cfg)
    N = G["N"]
    for it, ref in enumerate(G["steps"]):
        lab, unl = _batch()
        lab_k, _ = _batch()
        _, unl_k = _batch()
        ratios = ref["ratios"]  # random.uniform draws in the reference's order: unlabel_q first, then label_q
        out = O.run_step(student, teacher, opt, (lab, lab_k, unl, unl_k), cfg, ratios[:N], ratios[N:],
                         keep_rate=0.0 if it == 0 else None)
        for k, v in ref["losses"].items():
            assert abs(out[k] - v) <= 2e-5 * max(abs(v), 1e-6), (it, k, out[k], v)
        for name, model in (("student", student), ("teacher", teacher)):
            sd = model.ref_state_dict()
            assert set(sd) == set(ref[name])
            for k, v in ref[name].items():
                mine = sd[k].detach().reshape(-1)[_sample_idx(sd[k].numel())]
                err = float((mine - v).abs().max())
                assert err <= 2e-6 + 2e-5 * float(v.abs().max()), (it, name, k, err)
    # step 0 copied the student into the teacher before the student's update; step 1 moved it by (1 - 0.9996)
    k = "roi_heads.box_predictor.cls_score.weight"
    t0, t1 = G["steps"][0]["teacher"][k], G["steps"][1]["teacher"][k]
    assert not torch.equal(t0, t1)


GB = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_burnin_golden.pt"), weights_only=False)


def _burnin_batch():
    H, W = GB["H"], GB["W"]
    out = []
    for tag in ("q", "k"):
        out.append([{"image": im.clone(), "height": H, "width": W,
                     "instances": O.OInst((H, W), gt_boxes=O.OBoxes(b.clone()), gt_classes=c.clone())}
                    for im, b, c in zip(GB[f"lab_{tag}_images"], GB[f"gt_boxes_{tag}"], GB[f"gt_classes_{tag}"])])
    return out


def test_two_burn_in_steps_match_the_reference_trainer():
    """Source-only iterations (iter < BURN_UP_STEP, pt/engine/trainer.py:274-290): fixture made by
    oracle/make_golden_burnin.py from the reference's own PTrainer.run_step."""
    cfg = O.OracleCfg(num_classes=GB["K"], base_lr=GB["lr"])
    student = O.OracleRCNN(cfg, seed=GB["seed"])
    student.sampler = _Sampler(GB["prio"])
    opt = O.make_optimizer(student, cfg)
    for it, ref in enumerate(GB["steps"]):
        lab_q, lab_k = _burnin_batch()
        assert len(ref["ratios"]) == 2 * GB["N"]  # every image of q + k is resized, q first
        out = O.run_step_burn_in(student, opt, (lab_q, lab_k), cfg, ref["ratios"])
        assert set(ref["losses"]) == {"loss_cls", "loss_box_reg", "loss_rpn_cls", "loss_rpn_loc"}
        for k, v in ref["losses"].items():
            assert abs(out[k] - v) <= 2e-5 * max(abs(v), 1e-6), (it, k, out[k], v)
        sd = student.ref_state_dict()
        for k, v in ref["student"].items():
            mine = sd[k].detach().reshape(-1)[_sample_idx(sd[k].numel())]
            err = float((mine - v).abs().max())
            assert err <= 2e-6 + 2e-5 * float(v.abs().max()), (it, k, err)


def test_three_steps_with_trainer_hyper_parameters_off_their_defaults():
    """tests/golden/pt_reference_step_oddcfg_golden.pt (oracle/make_golden_step_oddcfg.py): the reference's own
    PTrainer.run_step for three iterations with loss weights 0.5 / 2.0, EMA keep rate 0.99, TEACHER_UPDATE_ITER 2
    (copy, skip, EMA), momentum 0.8, weight decay 5e-4, lr 0.004."""
    G3 = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_step_oddcfg_golden.pt"), weights_only=False)
    t = G3["trainer_cfg"]
    cfg = O.OracleCfg(num_classes=G3["K"], base_lr=t["base_lr"], momentum=t["momentum"], weight_decay=t["weight_decay"],
                      ema_keep_rate=t["ema_keep_rate"], source_loss_weight=t["source_loss_weight"],
                      target_unsup_loss_weight=t["target_unsup_loss_weight"])
    student = O.OracleRCNN(cfg, seed=G3["seed"])
    teacher = O.OracleRCNN(cfg, seed=G3["teacher_seed"])
    student.sampler = _Sampler(G3["prio"])
    opt = O.make_optimizer(student, cfg)
    N, H, W = G3["N"], G3["H"], G3["W"]

    def batch():
        lab = [{"image": im.clone(), "height": H, "width": W,
                "instances": O.OInst((H, W), gt_boxes=O.OBoxes(b.clone()), gt_classes=c.clone())}
               for im, b, c in zip(G3["lab_images"], G3["gt_boxes"], G3["gt_classes"])]
        unl = [{"image": im.clone(), "height": H, "width": W} for im in G3["unl_images"]]
        return lab, unl

    for it, ref in enumerate(G3["steps"]):
        lab, unl = batch()
        lab_k, _ = batch()
        _, unl_k = batch()
        ratios = ref["ratios"]
        out = O.run_step(student, teacher, opt, (lab, lab_k, unl, unl_k), cfg, ratios[:N], ratios[N:],
                         keep_rate=0.0 if it == 0 else None, update_teacher=it % t["teacher_update_iter"] == 0)
        for k, v in ref["losses"].items():
            assert abs(out[k] - v) <= 2e-5 * max(abs(v), 1e-6), (it, k, out[k], v)
        for name, model in (("student", student), ("teacher", teacher)):
            sd = model.ref_state_dict()
            for k, v in ref[name].items():
                mine = sd[k].detach().reshape(-1)[_sample_idx(sd[k].numel())]
                err = float((mine - v).abs().max())
                assert err <= 2e-6 + 2e-5 * float(v.abs().max()), (it, name, k, err)
