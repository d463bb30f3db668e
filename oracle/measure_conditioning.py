"""TEST INFRASTRUCTURE. Measures how well-conditioned the quantities of the reference trainer-step fixtures are.

A Probabilistic Teacher iteration contains discrete decisions (NMS keep masks, top-k, IoU-threshold matching with the
`iou == max` low-quality rule of d2's Matcher, priority-ordered sampling). On the small synthetic images of the step
fixtures many anchors are LARGER than the image, so all anchors of one shape that contain a (pseudo) box have the same
IoU = area(box) / area(anchor) in exact arithmetic; which of them pass the `==` test depends on the last bit of the
box coordinates. Quantities downstream of such a decision cannot agree to 1e-3 between ANY two fp32 implementations
(the reference on another GPU / cuDNN version included). This script quantifies that with the oracle, which is
pinned to the reference's own trainer on these fixtures (tests/test_oracle_golden_step.py): every step is replayed
once exactly and several times with the initial weights perturbed by relative Gaussian noise of 2e-6 (about 16 fp32
ulps; the anchor parameters, which both implementations hold exactly, are left alone), and the largest deviation of
every loss and of every sampled per-tensor parameter update is recorded.

    python oracle/measure_conditioning.py            # writes tests/golden/step_conditioning.json

The GPU trainer tests in the f16x3 precision (tests/test_trainer_step_gpu.py, tests/test_zz_next_rows_gpu.py) hold
every quantity to max(1e-3, 4 x its recorded conditioning): 1e-3 wherever the reference's own algorithm is
well-conditioned, and no tighter than the algorithm itself elsewhere."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pt_oracle as O  # noqa: E402

NOISE = 2e-6
SEEDS = (1, 2, 3)
GOLD = os.path.join(ROOT, "tests", "golden")


class _Sampler:
    def __init__(self, pr):
        self.pr = pr

    def prio(self, tag, n):
        grp, which = tag[0].split("_")
        return self.pr[grp][0 if which == "pos" else 1][tag[1]][:n]


def _idx(numel, n=64):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def _samples(model):
    return {k: v.detach().reshape(-1)[_idx(v.numel())].clone() for k, v in model.ref_state_dict().items()}


def _perturbed(sd, seed):
    if seed is None:
        return sd
    g = torch.Generator().manual_seed(seed)
    return {k: (v.detach() if "anchor_0" in k else v.detach() * (1 + NOISE * torch.randn(v.shape, generator=g)))
            for k, v in sd.items()}


def _replay(G, kind, seed):
    """Returns per step: (losses, student samples, teacher samples or None)."""
    t = G.get("trainer_cfg")
    kw = dict(num_classes=G["K"])
    if t:
        kw.update(base_lr=t["base_lr"], momentum=t["momentum"], weight_decay=t["weight_decay"],
                  ema_keep_rate=t["ema_keep_rate"], source_loss_weight=t["source_loss_weight"],
                  target_unsup_loss_weight=t["target_unsup_loss_weight"])
    else:
        kw.update(base_lr=G["lr"])
    cfg = O.OracleCfg(**kw)
    student = O.OracleRCNN(cfg, seed=G["seed"])
    student.load_ref_state_dict(_perturbed(student.ref_state_dict(), seed))
    student.sampler = _Sampler(G["prio"])
    opt = O.make_optimizer(student, cfg)
    H, W, N = G["H"], G["W"], G["N"]
    out = []
    if kind == "burnin":
        for ref in G["steps"]:
            views = []
            for tag in ("q", "k"):
                views.append([{"image": im.clone(), "height": H, "width": W,
                               "instances": O.OInst((H, W), gt_boxes=O.OBoxes(b.clone()), gt_classes=c.clone())}
                              for im, b, c in zip(G[f"lab_{tag}_images"], G[f"gt_boxes_{tag}"], G[f"gt_classes_{tag}"])])
            losses = O.run_step_burn_in(student, opt, views, cfg, ref["ratios"])
            out.append((losses, _samples(student), None))
        return out
    teacher = O.OracleRCNN(cfg, seed=G["teacher_seed"])

    def batch():
        lab = [{"image": im.clone(), "height": H, "width": W,
                "instances": O.OInst((H, W), gt_boxes=O.OBoxes(b.clone()), gt_classes=c.clone())}
               for im, b, c in zip(G["lab_images"], G["gt_boxes"], G["gt_classes"])]
        unl = [{"image": im.clone(), "height": H, "width": W} for im in G["unl_images"]]
        return lab, unl
    for it, ref in enumerate(G["steps"]):
        lab, unl = batch()
        lab_k, _ = batch()
        _, unl_k = batch()
        r = ref["ratios"]
        kws = dict(keep_rate=0.0 if it == 0 else None)
        if t:
            kws["update_teacher"] = it % t["teacher_update_iter"] == 0
        losses = O.run_step(student, teacher, opt, (lab, lab_k, unl, unl_k), cfg, r[:N], r[N:], **kws)
        out.append((losses, _samples(student), _samples(teacher)))
    return out


def measure(name, kind):
    G = torch.load(os.path.join(GOLD, name), weights_only=False)
    base = _replay(G, kind, None)
    init = {k: v.detach().reshape(-1)[_idx(v.numel())].clone()
            for k, v in O.OracleRCNN(O.OracleCfg(num_classes=G["K"]), seed=G["seed"]).ref_state_dict().items()}
    steps = []
    for it in range(len(base)):
        steps.append({"losses": {k: 0.0 for k in G["steps"][it]["losses"]}, "update": {}, "teacher": {}})
    for seed in SEEDS:
        pert = _replay(G, kind, seed)
        for it, ((l0, s0, t0), (l1, s1, t1)) in enumerate(zip(base, pert)):
            rec = steps[it]
            for k in rec["losses"]:
                rec["losses"][k] = max(rec["losses"][k], abs(l0[k] - l1[k]) / max(abs(l0[k]), 1e-6))
            prev0 = init if it == 0 else base[it - 1][1]
            prev1 = init if it == 0 else pert[it - 1][1]   # (initial perturbation itself: 2e-6 of the weights)
            for k in s0:
                du0, du1 = s0[k] - prev0[k], s1[k] - prev1[k]
                m = float(du0.abs().max())
                if m > 0:
                    rec["update"][k] = max(rec["update"].get(k, 0.0), float((du0 - du1).abs().max()) / m)
            if t0 is not None:
                for k in t0:
                    m = float(t0[k].abs().max())
                    if m > 0:
                        rec["teacher"][k] = max(rec["teacher"].get(k, 0.0), float((t0[k] - t1[k]).abs().max()) / m)
    return steps


def main():
    out = {"noise": NOISE, "seeds": list(SEEDS),
           "how": "max over seeds of |x(perturbed) - x| / max|x| with the oracle; see oracle/measure_conditioning.py"}
    for name, kind in (("pt_reference_step_golden.pt", "step"), ("pt_reference_step_oddcfg_golden.pt", "step"),
                       ("pt_reference_burnin_golden.pt", "burnin")):
        out[name] = measure(name, kind)
        for it, rec in enumerate(out[name]):
            print(name, "step", it, {k: float(f"{v:.1e}") for k, v in rec["losses"].items()},
                  "worst update", max(rec["update"].values()) if rec["update"] else None)
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", GOLD), "step_conditioning.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", dst)


if __name__ == "__main__":
    main()
