"""Repeats the supervised f16x3 forward of the small parity fixture and reports every quantity that changes from
run to run (proposal sets, losses): a racy kernel in the proposal path shows up here."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_parity_x3_gpu as T

cuda = torch.device("cuda:0")
H, W, K = 192, 272, 8
O, model, om = T._pair(cuda, K, "DifferentiableAnchorGenerator", 3)
lab = O.synthetic_batch(2, H, W, K, 1)
pr = T._prios(cuda, 2, H, W, 7)
model.prio_override = pr
first = None
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
nbad = 0
with torch.no_grad():
    for i in range(reps):
        lg, _, _, _ = model(T._to_inst(lab), branch="supervised")
        p = model._last_ctx["props"]
        cur = {"boxes": p["boxes"].clone(), "scores": p["scores"].clone(), "count": p["count"].clone(),
               "losses": {k: float(v) for k, v in lg.items()}}
        if first is None:
            first = cur
            print("first", cur["losses"], cur["count"].tolist(), flush=True)
            continue
        same_props = torch.equal(cur["boxes"], first["boxes"]) and torch.equal(cur["count"], first["count"])
        dl = {k: abs(cur["losses"][k] - first["losses"][k]) / max(abs(first["losses"][k]), 1e-9) for k in cur["losses"]}
        if not same_props or max(dl.values()) > 1e-5:
            nbad += 1
            print("rep", i, "props equal", same_props, "counts", cur["count"].tolist(), {k: f"{v:.2e}" for k, v in dl.items()}, flush=True)
print("changed repeats:", nbad, "of", reps - 1)
