"""Dev diagnostic (GPU): where do the f16x3 parameter gradients of the supervised branch differ from the oracle's?
Prints every tensor's relative error, the agreement of the sampled roi sets and of the head outputs / unit gradients."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import test_x3_backward_gpu as T  # noqa: E402

cuda = torch.device("cuda:0")
H, W, K = T.H, T.W, T.K
branch = sys.argv[1] if len(sys.argv) > 1 else "supervised"
O, model, om = T._setup(cuda)
g = torch.Generator().manual_seed(7)
R = (H // 16) * (W // 16) * 9
pr = {"rpn": (torch.rand(2, R, generator=g).to(cuda), torch.rand(2, R, generator=g).to(cuda)),
      "roi": (torch.rand(2, 2016, generator=g).to(cuda), torch.rand(2, 2016, generator=g).to(cuda))}
model.prio_override = pr
om.sampler = T._Sampler(pr)
model.zero_grad()
trace = {}
if branch == "supervised":
    lab = O.synthetic_batch(2, H, W, K, 1)
    lg, _, _, _ = model(T._to_inst(lab), branch="supervised")
    lo, _, _, _ = om(lab, branch="supervised", proposals_override=T._oracle_props(O, model, (H, W)), trace=trace)
else:
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    unl = O.synthetic_batch(2, H, W, K, 2, labelled=False)
    with torch.no_grad():
        _, _, roih, _ = om(unl, branch="unsup_data_weak")
    pseudo = [O.OInst(r.image_size, pseudo_boxes=O.OBoxes(r.pred_boxes.tensor), scores_logists=r.scores_logists,
                      boxes_sigma=r.boxes_sigma) for r in roih]
    unl_o = [dict(d, instances=p) for d, p in zip(unl, pseudo)]
    unl_g = [dict(d, instances=FreeInstances(p.image_size, pseudo_boxes=Boxes(p.pseudo_boxes.tensor.to(cuda)),
                                             scores_logists=p.scores_logists.to(cuda), boxes_sigma=p.boxes_sigma.to(cuda)))
             for d, p in zip(unl, pseudo)]
    lg, _, _, _ = model(unl_g, branch="unsupervised", danchor=True)
    lo, _, _, _ = om(unl_o, branch="unsupervised", danchor=True, proposals_override=T._oracle_props(O, model, (H, W)), trace=trace)
print({k: (float(lg[k]), float(lo[k])) for k in lo})
ctx = model._last_ctx["roi"]
haux = trace["roi"]
haux["scores"].retain_grad()
haux["deltas"].retain_grad()
sum(lg.values()).backward()
sum(lo.values()).backward()
torch.cuda.synchronize()
# sampled sets
counts = ctx["counts"].tolist()
cap = ctx["cap"]
rows = []
for n, c in enumerate(counts):
    rows += list(range(n * cap, n * cap + min(c, cap)))
rows = torch.tensor(rows, device=cuda)
props_o = torch.cat([O._bt(p.proposal_boxes) for p in haux["proposals"]])
print("sampled rois:", len(rows), "oracle", len(props_o), "max box diff",
      float((ctx["rois"].view(-1, 4)[rows].cpu() - props_o).abs().max()) if len(rows) == len(props_o) else "n/a")
if branch == "supervised":
    cls_o = torch.cat([p.gt_classes for p in haux["proposals"]])
    print("label mismatches:", int((ctx["sel"]["gt_classes"].view(-1)[rows].cpu() != cls_o).sum()))
print("scores rel", T._rel(ctx["scores"][rows], haux["scores"]), "deltas rel", T._rel(ctx["deltas"][rows], haux["deltas"]))
print("dscores rel", T._rel(ctx["dscores"][rows], haux["scores"].grad), "ddeltas rel",
      T._rel(ctx["ddeltas"][rows], haux["deltas"].grad))
gr = T._grad_errors(model, om)
for name, (r, m) in sorted(gr.items(), key=lambda kv: -kv[1][0]):
    print(f"{r:.2e}  max|g| {m:.3e}  {name}")
