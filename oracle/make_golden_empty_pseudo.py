"""Generates tests/golden/pt_reference_empty_pseudo_golden.pt: the unsupervised branch of the REFERENCE'S OWN MODEL
CLASSES (`GuassianGeneralizedRCNN.forward(branch="unsupervised", danchor=True)`) when the teacher delivered NO pseudo
label for one image of the batch, and for none at all -- the "empty input" edge of rows a-6 / a-8 / a-11 / a-15 --
plus the supervised branch on a batch in which no image has any ground-truth box.
The reference returns finite losses in the first case and (loss_cls, loss_box_reg) = NaN (a mean over zero rois),
(loss_rpn_cls, loss_rpn_loc) = 0 in the second.

    python oracle/make_golden_empty_pseudo.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import make_golden_model as M  # noqa: E402

O, d2shim_model = M.O, M.d2shim_model
from probabilisticteacher_b200.config import c2f_config  # noqa: E402

H, W, K, N, WEIGHT_SEED, UNL_SEED, PRIO_SEED = 128, 160, 8, 2, 11, 50, 99


def main():
    cfg = c2f_config()
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K), seed=WEIGHT_SEED).ref_state_dict()
    model = M.build_reference_model(cfg, {k: v.detach().clone() for k, v in sd.items()})
    model.train()
    unl = O.synthetic_batch(N, H, W, K, UNL_SEED, labelled=False)
    g = torch.Generator().manual_seed(PRIO_SEED)
    R, L = (H // 16) * (W // 16) * 9, 2000 + 16
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}
    d2shim_model.PRIO.provider = lambda tag, n: prio[tag[0].split("_")[0]][0 if tag[0].endswith("pos") else 1][tag[1]][:n]
    with torch.no_grad():
        d2shim_model.PRIO.reset()
        _, _, roih, _ = model(M.to_ref(unl), branch="unsup_data_weak")
    out = dict(H=H, W=W, K=K, N=N, weight_seed=WEIGHT_SEED, unl_seed=UNL_SEED, prio_seed=PRIO_SEED,
               teacher_roih=[dict(pred_boxes=p.pred_boxes.tensor.clone(), scores_logists=p.scores_logists.clone(),
                                  boxes_sigma=p.boxes_sigma.clone()) for p in roih], cases={})
    for case, keep in (("second_image_empty", [None, 0]), ("all_empty", [0, 0])):
        q = []
        for d, p, n in zip(M.to_ref(unl), roih, keep):
            n = len(p.pred_boxes) if n is None else n
            q.append(dict(d, instances=M.FreeInstances(p.image_size, pseudo_boxes=M.Boxes(p.pred_boxes.tensor[:n].clone()),
                                                       scores_logists=p.scores_logists[:n].clone(),
                                                       boxes_sigma=p.boxes_sigma[:n].clone())))
        with torch.no_grad():
            d2shim_model.PRIO.reset()
            losses, _, _, _ = model(q, branch="unsupervised", danchor=True)
        out["cases"][case] = dict(keep=keep, losses={k: float(v) for k, v in losses.items()})
        print(case, out["cases"][case]["losses"])
    # the unsupervised branch under other UNSUPNET settings than train.sh's (configs/pt/final_c2f.yaml itself says
    # TAU [0.25, 0.25]; EFL off; other EFL_LAMBDA exponents), full pseudo labels: exercises the `efl` / `tau` /
    # `lambda` arguments of the loss kernels
    out["unsupnet_variants"] = []
    for efl, tau, lam in ((False, [0.25, 0.25], [0.5, 0.5]), (True, [0.5, 0.25], [1.0, 2.0])):
        vcfg = c2f_config()
        vcfg.UNSUPNET.EFL, vcfg.UNSUPNET.TAU, vcfg.UNSUPNET.EFL_LAMBDA = efl, tau, lam
        vmodel = M.build_reference_model(vcfg, {k: v.detach().clone() for k, v in sd.items()})
        vmodel.train()
        q = [dict(d, instances=M.FreeInstances(p.image_size, pseudo_boxes=M.Boxes(p.pred_boxes.tensor.clone()),
                                               scores_logists=p.scores_logists.clone(), boxes_sigma=p.boxes_sigma.clone()))
             for d, p in zip(M.to_ref(unl), roih)]
        with torch.no_grad():
            d2shim_model.PRIO.reset()
            losses, _, _, _ = vmodel(q, branch="unsupervised", danchor=True)
        out["unsupnet_variants"].append(dict(efl=efl, tau=tau, efl_lambda=lam, losses={k: float(v) for k, v in losses.items()}))
        print("variant", efl, tau, lam, out["unsupnet_variants"][-1]["losses"])
    # supervised branch when NO image of the batch has ground truth (rows a-6 / a-7 / a-11 / a-14): every anchor and
    # every proposal is background, both regression losses are (minus) zero
    lab = O.synthetic_batch(N, H, W, K, UNL_SEED + 10, boxes_per_image=0)
    with torch.no_grad():
        d2shim_model.PRIO.reset()
        losses, _, _, _ = model(M.to_ref(lab), branch="supervised")
    out["supervised_no_gt"] = dict(lab_seed=UNL_SEED + 10, losses={k: float(v) for k, v in losses.items()})
    print("supervised_no_gt", out["supervised_no_gt"]["losses"])
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_empty_pseudo_golden.pt")
    torch.save(out, dst)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
