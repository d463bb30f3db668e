"""GPU: the B200 `PTrainer.run_step` against TWO STEPS OF THE REFERENCE'S OWN TRAINER
(tests/golden/pt_reference_step_golden.pt, made by oracle/make_golden_step.py from the unmodified
pt/engine/trainer.py + pt/modeling classes): same weights, images, `resize` ratios and sampling priorities.

The training path computes in fp16 operands / fp32 accumulation, so losses are compared at the fp16 tolerances of
tests/test_e2e_gpu.py; what must hold tightly is the trainer logic itself: the teacher after step 0 IS the initial
student (copy, trainer.py:293-295), the teacher after step 1 is the EMA of that copy and the updated student
(trainer.py:431-449), and the student's parameter update (clip + SGD with momentum and weight decay,
trainer.py:383-386,592-603) points where the reference's update points."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

import json

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_step_golden.pt"), weights_only=False)
# how well-conditioned each quantity of the fixture is under a 2e-6 relative perturbation of the weights, measured with
# the reference-pinned oracle (oracle/measure_conditioning.py): the f16x3 step is held to max(1e-3, 4 x conditioning)
COND = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "step_conditioning.json")))


def x3_tol(cond):
    return max(1e-3, 4.0 * cond)


def _sample_idx(numel, n=64):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


class _Ratios:
    def __init__(self, draws):
        self.draws = list(draws)

    def uniform(self, a, b):
        return self.draws.pop(0)


def _samples(model):
    return {k: v.detach().reshape(-1).cpu()[_sample_idx(v.numel())] for k, v in model.state_dict().items()}


@pytest.mark.parametrize("precision", ["f16x3", "f16"])
def test_two_steps_vs_reference_trainer(cuda, precision):
    """precision="f16x3" (fp32-equivalent forward + backward) is held to north_star's 1e-3 wherever the reference's
    own algorithm is that well-conditioned on this fixture, and to 4 x the measured conditioning elsewhere
    (tests/golden/step_conditioning.json: e.g. the unsupervised RPN losses move by 15 % when the weights are perturbed
    by 2e-6 -- exact IoU ties between same-shape anchors that contain a pseudo box, decided by d2 Matcher's
    `iou == max` test). Checked: the 8 losses of both steps and every sampled per-tensor parameter update."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    H, W, K = G["H"], G["W"], G["K"]
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.SOLVER.WARMUP_ITERS = 0      # the fixture's optimizer has the constant lr of the config (no scheduler hook)
    cfg.SOLVER.BASE_LR = G["lr"]

    def batch():
        lab = [{"image": im.clone(), "height": H, "width": W,
                "instances": FreeInstances((H, W), gt_boxes=Boxes(b.clone()), gt_classes=c.clone())}
               for im, b, c in zip(G["lab_images"], G["gt_boxes"], G["gt_classes"])]
        unl = [{"image": im.clone(), "height": H, "width": W} for im in G["unl_images"]]
        return lab, unl

    def loader():
        while True:
            lab, unl = batch()
            lab_k, _ = batch()
            _, unl_k = batch()
            yield lab, lab_k, unl, unl_k

    tr = PTrainer(cfg, loader(), device=cuda, seed=0, precision=precision)
    x3 = precision == "f16x3"
    ocfg = O.OracleCfg(num_classes=K)
    sd = {k: v.detach() for k, v in O.OracleRCNN(ocfg, seed=G["seed"]).ref_state_dict().items()}
    sd_t = {k: v.detach() for k, v in O.OracleRCNN(ocfg, seed=G["teacher_seed"]).ref_state_dict().items()}
    tr.model.load_state_dict(sd)
    tr.model_teacher.load_state_dict(sd_t)
    tr.model.prio_override = {k: (v[0].to(cuda), v[1].to(cuda)) for k, v in G["prio"].items()}
    init = {k: v.reshape(-1)[_sample_idx(v.numel())] for k, v in sd.items()}
    prev_student = init
    problems = []
    for it, ref in enumerate(G["steps"]):
        tr.rng = _Ratios(ref["ratios"])
        losses = tr.run_step()
        torch.cuda.synchronize()
        got = {k: float(v) for k, v in losses.items()}
        print("step", it, {k: (round(got[k], 4), round(v, 4)) for k, v in ref["losses"].items()})
        if x3:
            cond = COND["pt_reference_step_golden.pt"][it]
            for k, v in ref["losses"].items():
                tol = x3_tol(cond["losses"][k])
                print(f"      {k}: rel err {abs(got[k] - v) / max(abs(v), 1e-3):.2e} (tolerance {tol:.1e})")
                if not abs(got[k] - v) <= tol * max(abs(v), 1e-3):
                    problems.append((it, k, got[k], v, tol))
        elif it == 0:  # (step 1 starts from fp16-path weights: its losses drift with the discrete proposal selection)
            for k, v in ref["losses"].items():
                # supervised losses depend on this model's fp16 forward only; the unsupervised ones also on the
                # teacher's top-100 pseudo labels (near-tied scores at the synthetic initialisation)
                tol = (1e-2 if "rpn" in k else 5e-2) if k.endswith("_sup") else 0.3
                if not abs(got[k] - v) <= tol * max(abs(v), 1e-3):
                    problems.append((it, k, got[k], v))
        st, te = _samples(tr.model), _samples(tr.model_teacher)
        # ---- teacher: copy at step 0 (exact), EMA at step 1
        if it == 0:
            for k, v in init.items():
                assert torch.equal(te[k], v), k
        else:
            for k, v in ref["teacher"].items():
                err = float((te[k] - v).abs().max())
                assert err <= 1e-6 + 1e-5 * float(v.abs().max()), (k, err)
        # ---- student: direction and size of the update over all sampled parameters
        up_g = torch.cat([st[k] - prev_student[k] for k in sorted(st)])
        up_r = torch.cat([ref["student"][k] - (init[k] if it == 0 else G["steps"][0]["student"][k]) for k in sorted(st)])
        cos = float(torch.dot(up_g, up_r) / (up_g.norm() * up_r.norm()))
        ratio = float(up_g.norm() / up_r.norm())
        print("   update cosine", round(cos, 6), "norm ratio", round(ratio, 6))
        if x3:
            # per-tensor parameter update against the reference trainer's: max |du - du_ref| <= 1e-3 max |du_ref|
            prev_ref = init if it == 0 else G["steps"][0]["student"]
            worst = ("", 0.0, 0.0)
            for k in sorted(st):
                du, dr = st[k] - prev_student[k], ref["student"][k] - prev_ref[k]
                if float(dr.abs().max()) == 0.0:
                    assert float(du.abs().max()) == 0.0, k
                    continue
                e = float((du - dr).abs().max() / dr.abs().max())
                tol = x3_tol(cond["update"].get(k, 0.0))
                if e / tol > worst[1]:
                    worst = (k, e / tol, e)
                if e > tol:
                    problems.append(("update_x3", it, k, e, tol))
            print("   worst per-tensor update error / tolerance", worst)
            prev_student = st
            continue
        # measured: cosine 0.9998 / 0.9992, norm ratio 0.997 / 1.001 (steps 0 / 1)
        if not (cos > (0.995 if it == 0 else 0.99) and (0.97 if it == 0 else 0.95) < ratio < (1.03 if it == 0 else 1.05)):
            problems.append(("update", it, cos, ratio))
        frozen = [k for k in st if k.startswith("backbone.vgg_block1") or k.startswith("backbone.vgg_block2")]
        for k in frozen:  # FREEZE_AT = 2
            assert torch.equal(st[k], init[k]), k
        prev_student = st
    assert not problems, problems
