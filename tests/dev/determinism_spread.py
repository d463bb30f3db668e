"""Dev (GPU): run-to-run spread of one training-branch forward + backward from identical state (fp32 atomics in the
weight-gradient split-K, split-K fc1 and the ROIAlign backward accumulate in arrival order)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from probabilisticteacher_b200.config import c2f_config  # noqa: E402
from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model  # noqa: E402
from probabilisticteacher_b200.synthetic import synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (800, 1333)
for prec in ("f16x3", "f16"):
    model = build_model(c2f_config(), dev, precision=prec)
    model.init_synthetic(3)
    model.train()
    lab = synthetic_batch(2, H, W, 8, 1)
    g = torch.Generator().manual_seed(7)
    R = (H // 16) * (W // 16) * 9
    model.prio_override = {"rpn": (torch.rand(2, R, generator=g).to(dev), torch.rand(2, R, generator=g).to(dev)),
                           "roi": (torch.rand(2, 2016, generator=g).to(dev), torch.rand(2, 2016, generator=g).to(dev))}
    runs = []
    for _ in range(3):
        model.zero_grad()
        losses, _, _, _ = model(lab, branch="supervised")
        sum(losses.values()).backward()
        torch.cuda.synchronize()
        runs.append(({k: float(v) for k, v in losses.items()}, model.arena.grads.clone()))
    worst_l = max(abs(runs[0][0][k] - r[0][k]) / abs(runs[0][0][k]) for r in runs[1:] for k in runs[0][0])
    worst_g = 0.0
    worst_name = ""
    for name, v, gv, tr in model.arena.exposed_parameters():
        if not tr:
            continue
        o = model.arena.segments.get(name)
    a = model.arena
    for s in a.segments.values():
        if not s.trainable:
            continue
        lo = s.offset - a.trainable_start
        g0 = runs[0][1][lo:lo + s.numel]
        m = float(g0.abs().max())
        for r in runs[1:]:
            d = float((r[1][lo:lo + s.numel] - g0).abs().max())
            if m > 0 and d / m > worst_g:
                worst_g, worst_name = d / m, s.name
    print(f"{prec} {H}x{W}: 3 identical forward+backward passes: worst loss spread {worst_l:.2e} (relative), worst "
          f"gradient spread {worst_g:.2e} of the tensor's max ({worst_name})")
