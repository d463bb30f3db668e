"""Dev diagnostic (GPU): the box-head backward of the f16x3 precision, step by step against fp64 torch on the device's own
saved tensors, twice (determinism)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import test_x3_backward_gpu as T  # noqa: E402
from probabilisticteacher_b200 import ops  # noqa: E402

cuda = torch.device("cuda:0")
H, W, K = T.H, T.W, T.K
O, model, om = T._setup(cuda)
g = torch.Generator().manual_seed(7)
R = (H // 16) * (W // 16) * 9
pr = {"rpn": (torch.rand(2, R, generator=g).to(cuda), torch.rand(2, R, generator=g).to(cuda)),
      "roi": (torch.rand(2, 2016, generator=g).to(cuda), torch.rand(2, 2016, generator=g).to(cuda))}
model.prio_override = pr
lab = O.synthetic_batch(2, H, W, K, 1)
model.zero_grad()
lg, _, _, _ = model(T._to_inst(lab), branch="supervised")
ctx = model._last_ctx["roi"]
ar = model.arena
fc = ar.fc_dim
rows = ctx["x0"].shape[0]
seg = (ctx["counts"], ctx["cap"])
S = 1024.0
one = torch.ones(1, device=cuda)
live = torch.zeros(rows, dtype=torch.bool, device=cuda)
for n, c in enumerate(ctx["counts"].tolist()):
    live[n * ctx["cap"]:n * ctx["cap"] + min(c, ctx["cap"])] = True
rel = T._rel


def up(t3, k):
    return ops.split3_unpack(t3.contiguous(), k).double()


h2, h1 = up(ctx["h2"], fc), up(ctx["h1"], fc)
Wp = ar.view("roi_heads.box_predictor._heads.weight").double()      # [128][fc]
W2 = ar.view("roi_heads.box_head.fc2.weight").double()              # [fc][fc]
for trial in range(2):
    dpred3 = ops.pack_grad2_x3(ctx["dscores"], K + 1, ctx["ddeltas"], 8 * K, one, one, S, rows, 128).view(1, rows, 384)
    dpred = up(dpred3, 128)
    wd3, alpha = ar.dgrad_x3["pred"]
    dz2_3 = ops.gemm_tn_x3(dpred3, wd3, alpha, epi=ops.EPI_SPLIT3_MASK, aux=ctx["h2"].view(1, rows, 3 * fc), seg=seg)
    dz2 = up(dz2_3, fc)
    dz2_ref = (dpred @ Wp) * (h2 > 0)
    print("trial", trial, "dz2 rel (live rows)", rel(dz2[live], dz2_ref[live]), "dead rows max", float(dz2[~live].abs().max()),
          "ref dead max", float(dz2_ref[~live].abs().max()))
    bad = ((dz2 - dz2_ref).abs() > 1e-4 * dz2_ref.abs().max()) & live[:, None]
    print("   bad elements", int(bad.sum()), "rows", sorted(set(bad.nonzero()[:, 0].tolist()))[:10], "cols",
          sorted(set(bad.nonzero()[:, 1].tolist()))[:10])
    gw = torch.zeros(fc, fc, device=cuda)
    gb = torch.zeros(fc, device=cuda)
    ops.wgrad_x3(dz2_3, ctx["h1"].view(1, rows, 3 * fc), gw, m_total=fc, n_total=fc, scale=1.0 / S, bias_out=gb, seg=seg)
    gw_ref = (dz2.t() @ h1) / S
    gb_ref = dz2.sum(0) / S
    print("   wgrad rel", rel(gw, gw_ref), "bias rel", rel(gb, gb_ref), " (vs the device's own dz2)")
    gw_ref2 = ((dz2 * live[:, None]).t() @ h1) / S
    print("   wgrad rel using live rows of dz2 only", rel(gw, gw_ref2), "bias", rel(gb, (dz2 * live[:, None]).sum(0) / S))
    wd3, alpha = ar.dgrad_x3["fc2"]
    dz1_3 = ops.gemm_tn_x3(dz2_3, wd3, alpha, epi=ops.EPI_SPLIT3_MASK, aux=ctx["h1"].view(1, rows, 3 * fc), seg=seg)
    dz1 = up(dz1_3, fc)
    dz1_ref = (dz2 @ W2) * (h1 > 0)
    print("   dz1 rel (live rows)", rel(dz1[live], dz1_ref[live]))
    torch.cuda.synchronize()
