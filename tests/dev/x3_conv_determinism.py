"""Runs the same f16x3 conv repeatedly and compares the outputs bit for bit (a racy pipeline shows up here)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from probabilisticteacher_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (Cin, Cout, H, W, n) in [(64, 64, 800, 1333, 2), (64, 64, 96, 136, 4), (64, 128, 400, 666, 2), (128, 128, 48, 68, 4), (128, 256, 200, 333, 2)]:
    x = (torch.randn(n, Cin, H, W, generator=g).abs()).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cout)) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    xf = torch.zeros(n, H, W + 1, Cin, device=dev)
    xf[:, :, :W] = x.permute(0, 2, 3, 1)
    x3 = ops.FlatAct(ops.split3_pack(xf, Cin).view(n, H * (W + 1), 3 * Cin), H, W)
    w3 = ops.split3_pack(w.permute(0, 2, 3, 1).contiguous(), Cin, 1024.0, 1).view(Cout, -1)
    first = ops.conv3x3_x3(x3, w3, 1.0 / 1024.0, b).t.clone()
    bad = 0
    for i in range(300):
        y = ops.conv3x3_x3(x3, w3, 1.0 / 1024.0, b).t
        if not torch.equal(y, first):
            bad += 1
            d = (y.float() - first.float()).abs()
            print("  mismatch at rep", i, "n_diff", int((d > 0).sum()), "max", float(d.max()), flush=True)
            if bad > 5:
                break
    print(f"conv {Cin}->{Cout} {H}x{W} x{n}: {bad} mismatching repeats", flush=True)
