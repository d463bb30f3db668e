import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import pt_oracle as O
from probabilisticteacher_b200 import ops
from probabilisticteacher_b200._lib import lib
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
H, W = 50, 83
cell = O.differentiable_cell_anchors(torch.tensor(O.OracleCfg().anchor_wh))
anchors = O.grid_anchors(cell, H, W, 16, 0.0)
N, k = 2, 12000
boxes = []
for n in range(N):
    idx = torch.randperm(anchors.shape[0], generator=g)[:k]
    b = anchors[idx] + torch.randn(k, 4, generator=g) * 4
    b = O.clip_boxes(b, (800, 1333))
    boxes.append(b)
boxes = torch.stack(boxes).to(dev)
order = torch.arange(k, dtype=torch.int32).repeat(N, 1).to(dev)
counts = torch.tensor([k, k], dtype=torch.int32, device=dev)
for _ in range(2):
    ki, kc = ops.nms(boxes, order, counts, 0.7, 2000)
torch.cuda.synchronize()
print("kept", kc.tolist())
L = lib()
has_dbg = hasattr(L, "ptb200_nms_debug_read")
if has_dbg:
    buf = (ctypes.c_longlong * 8)()
    L.ptb200_nms_debug_read(buf)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ki, kc = ops.nms(boxes, order, counts, 0.7, 2000)
e1.record(); torch.cuda.synchronize()
print("nms (bitmask+scan) ms:", e0.elapsed_time(e1) / 5)
if has_dbg:
    L.ptb200_nms_debug_read(buf)
    v = list(buf)
    steps = max(v[7], 1)
    names = ["prefetch issue", "cp.async wait", "sync A", "resolve", "sync B", "OR phase", "sync C"]
    for nm, c in zip(names, v[:7]):
        print(f"  {nm:16s} {c / steps:8.0f} cycles/step")
    print("  steps", steps / 10)
