"""Times the f16x3 forward convs of VGG blocks 1-2 (N = 64 / 128: the row-window kernel) at full size.
PTB200_X3_ROWWIN=0 runs the per-tap kernel for an A/B; PTB200_X3_RW_CHUNK sets k-iterations per promotion."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from probabilisticteacher_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for (Cin, Cout, H, W, n) in [(64, 64, 800, 1333, 2), (64, 128, 400, 666, 2), (128, 128, 400, 666, 2), (128, 256, 200, 333, 2), (256, 256, 200, 333, 2),
                           (256, 512, 100, 167, 2), (512, 512, 100, 167, 2), (512, 512, 50, 84, 4)]:
    x = (torch.randn(n, Cin, H, W, generator=g).abs()).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (2.0 / (9 * Cout)) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    xf = torch.zeros(n, H, W + 1, Cin, device=dev)
    xf[:, :, :W] = x.permute(0, 2, 3, 1)
    x3 = ops.FlatAct(ops.split3_pack(xf, Cin).view(n, H * (W + 1), 3 * Cin), H, W)
    w3 = ops.split3_pack(w.permute(0, 2, 3, 1).contiguous(), Cin, 1024.0, 1).view(Cout, -1)
    y3 = ops.conv3x3_x3(x3, w3, 1.0 / 1024.0, b)
    y = ops.split3_unpack(y3.t, Cout).view(n, H, W + 1, Cout)[:, :64, :W]
    ref = F.relu(F.conv2d(x[:, :, :65].double(), w.double(), b.double(), padding=1)).permute(0, 2, 3, 1)[:, :64]
    err = float((y.double() - ref).norm() / ref.norm())
    for _ in range(3):
        ops.conv3x3_x3(x3, w3, 1.0 / 1024.0, b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv3x3_x3(x3, w3, 1.0 / 1024.0, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * n * H * (W + 1) * Cout * 9 * 3 * Cin
    print(f"conv {Cin}->{Cout} {H}x{W} x{n}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s (x3 flops)  rel err {err:.2e}", flush=True)
