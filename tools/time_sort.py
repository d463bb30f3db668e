"""Times the segmented radix sort on score-like keys (RPN: 4 x 37350 keys; roi filter: 2 x 16000)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticteacher_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
for segs, n in ((4, 37350), (2, 37350), (2, 16000)):
    f = (torch.randn(segs, n, generator=g) * 0.05).to(dev)
    bits = f.view(torch.int32)
    keys0 = torch.where(bits < 0, ~bits, bits | (-2 ** 31)).contiguous()  # monotonic float -> uint32 order
    vals0 = torch.arange(n, dtype=torch.int32, device=dev).repeat(segs, 1)
    ts = []
    for it in range(6):
        keys, vals = keys0.clone(), vals0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.segmented_sort(keys, vals)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    # check against torch (unsigned order == order of the float values)
    ref = torch.sort(f, dim=1, stable=True)
    ok = torch.equal(vals.long(), ref.indices)
    print(f"segments={segs} n={n}: {sorted(ts)[len(ts)//2]*1e3:.1f} us  matches torch.sort(stable): {ok}")
