"""End-to-end comparison of the B200 path with the CPU oracle on a small synthetic batch (dev tool;
the pytest version lives in tests/test_e2e_gpu.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pt_oracle as O
from probabilisticteacher_b200.config import c2f_config
from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
from probabilisticteacher_b200.structures import Boxes, FreeInstances
from probabilisticteacher_b200 import ops


class FixedSampler:
    def __init__(self, prios):
        self.p = prios

    def prio(self, tag, n):
        kind, i = tag
        group, which = kind.split("_")
        pp, pn = self.p[group]
        return (pp if which == "pos" else pn)[i].cpu()


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def to_inst(batch):
    out = []
    for d in batch:
        nd = dict(d)
        if "instances" in d:
            i = d["instances"]
            nd["instances"] = FreeInstances(i.image_size, gt_boxes=Boxes(i.gt_boxes.tensor.clone()), gt_classes=i.gt_classes.clone())
        out.append(nd)
    return out


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    cfg = c2f_config()
    H, W = int(os.environ.get("H", 192)), int(os.environ.get("W", 272))
    student = build_model(cfg, dev)
    sd = student.init_synthetic(seed=3)
    ocfg = O.OracleCfg()
    om = O.OracleRCNN(ocfg, seed=0)
    om.load_ref_state_dict(sd)
    # ---- state_dict round trip
    sd2 = student.state_dict()
    for k in sd:
        assert torch.equal(sd2[k].cpu(), sd[k]), k
    print("state_dict round trip ok", len(sd))
    student.train()

    lab = O.synthetic_batch(2, H, W, 8, 1)
    unl = O.synthetic_batch(2, H, W, 8, 2, labelled=False)

    # ---- teacher pass
    tr = {}
    with torch.no_grad():
        _, props_o, roih_o, _ = om(unl, branch="unsup_data_weak", trace=tr)
        _, props_g, roih_g, _ = student(unl, branch="unsup_data_weak")
    torch.cuda.synchronize()
    feat_g = ops.from_flat(student._last_ctx["feat"]) if hasattr(student, "_last_ctx") else None
    print("flag", int(student.proposal_generator.nonfinite_flag))
    for n in range(2):
        pg = props_g[n].trim()
        po = props_o[n]
        print(f"img{n}: proposals gpu {len(pg)} oracle {len(po.proposal_boxes)}")
        k = min(len(pg), len(po.proposal_boxes), 50)
        print("   top boxes maxabs diff", float((pg.proposal_boxes.tensor[:k].cpu() - po.proposal_boxes.tensor[:k]).abs().max()))
        rg = roih_g[n].trim()
        ro = roih_o[n]
        print(f"   dets gpu {len(rg)} oracle {len(ro.scores)}; scores top5 gpu {rg.scores[:5].tolist()} oracle {ro.scores[:5].tolist()}")
        print("   classes gpu", rg.pred_classes[:10].tolist(), "oracle", ro.pred_classes[:10].tolist())

    # ---- supervised pass with injected sampling priorities
    N = 2
    Hf, Wf = H // 16, W // 16
    R = Hf * Wf * 9
    g = torch.Generator().manual_seed(7)
    L = 2000 + 16
    prios = {"rpn": (torch.rand(N, R, generator=g).to(dev), torch.rand(N, R, generator=g).to(dev)),
             "roi": (torch.rand(N, L, generator=g).to(dev), torch.rand(N, L, generator=g).to(dev))}
    student.prio_override = prios
    om.sampler = FixedSampler(prios)
    student.zero_grad()
    losses_g, _, _, _ = student(to_inst(lab), branch="supervised")
    tr = {}
    losses_o, _, _, _ = om(lab, branch="supervised", trace=tr)
    print("sup losses gpu   ", {k: float(v) for k, v in losses_g.items()})
    print("sup losses oracle", {k: float(v) for k, v in losses_o.items()})
    feat_g = ops.from_flat(student._last_ctx["feat"])
    print("feature rel err", rel(feat_g, tr["features"]))
    lg = student._last_ctx["rpn"]["logits"].view(N, Hf, Wf + 1, 9)[:, :, :Wf].reshape(N, -1)
    print("rpn logits rel err", rel(lg, tr["rpn"]["logits"]))
    sum(losses_g.values()).backward()
    for p in om.parameters():
        p.grad = None
    sum(losses_o.values()).backward()
    torch.cuda.synchronize()
    gsd = {}
    og = {k.replace("__", "."): v.grad for k, v in om.named_parameters()}
    worst = []
    kinds = {s.name: s.kind for s in student.arena.segments.values()}
    for (name, v, gv, trainable) in student.arena.exposed_parameters():
        if not trainable:
            continue
        kind = kinds.get(name, "mat")
        gr = student.arena._to_ref(kind, gv, student.arena.C, 7).reshape(og[name].shape)
        r = rel(gr, og[name])
        worst.append((r, name, float(og[name].abs().max())))
    for r, name, m in sorted(worst, reverse=True)[:40]:
        print(f"  grad rel err {r:.4f}  {name}  (max |g| {m:.3e})")

    # ---- unsupervised pass using oracle pseudo labels on both sides
    pseudo_o = [O.OInst(r.image_size, pseudo_boxes=O.OBoxes(r.pred_boxes.tensor), scores_logists=r.scores_logists,
                        boxes_sigma=r.boxes_sigma) for r in roih_o]
    unl_o = [dict(d, instances=p) for d, p in zip(unl, pseudo_o)]
    unl_g = [dict(d, instances=FreeInstances(p.image_size, pseudo_boxes=Boxes(p.pseudo_boxes.tensor.to(dev)),
                                             scores_logists=p.scores_logists.to(dev), boxes_sigma=p.boxes_sigma.to(dev)))
             for d, p in zip(unl, pseudo_o)]
    student.zero_grad()
    lg_, _, _, _ = student(unl_g, branch="unsupervised", danchor=True)
    lo_, _, _, _ = om(unl_o, branch="unsupervised", danchor=True)
    print("unsup losses gpu   ", {k: float(v) for k, v in lg_.items()})
    print("unsup losses oracle", {k: float(v) for k, v in lo_.items()})
    sum(lg_.values()).backward()
    for p in om.parameters():
        p.grad = None
    sum(lo_.values()).backward()
    og = {k.replace("__", "."): v.grad for k, v in om.named_parameters()}
    worst = []
    for (name, v, gv, trainable) in student.arena.exposed_parameters():
        if not trainable:
            continue
        kind = kinds.get(name, "mat")
        gr = student.arena._to_ref(kind, gv, student.arena.C, 7).reshape(og[name].shape)
        worst.append((rel(gr, og[name]), name, float(og[name].abs().max())))
    for r, name, m in sorted(worst, reverse=True)[:40]:
        print(f"  grad rel err {r:.4f}  {name}  (max |g| {m:.3e})")
    print("anchor grad gpu", student.arena.gview("proposal_generator.anchor_generator.anchor_0").flatten().tolist())
    print("anchor grad ora", og["proposal_generator.anchor_generator.anchor_0"].flatten().tolist())


if __name__ == "__main__":
    main()
