"""Dev: first timing of the f16x3 training step at full size (graph + concurrent), next to f16."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from probabilisticteacher_b200.config import c2f_config  # noqa: E402
from probabilisticteacher_b200.engine.trainer import PTrainer  # noqa: E402

dev = torch.device("cuda:0")
cfg = c2f_config()
cfg.UNSUPNET.BURN_UP_STEP = 0
pool = bench.synthetic_pool(2, 2, 800, 1333, 8, 1234, device=dev)
for prec in sys.argv[1:] or ["f16x3", "f16"]:
    for graph in (False, True):
        tr = PTrainer(cfg, bench.cycle(pool), device=dev, seed=0, use_cuda_graph=graph, concurrent=graph, precision=prec)
        for _ in range(8):
            tr.step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 20
        for _ in range(n):
            losses = tr.step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n * 1e3
        print(prec, "graph" if graph else "eager", f"{dt:.2f} ms/step", {k: round(float(v), 4) for k, v in losses.items()},
              f"mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
        del tr
        torch.cuda.empty_cache()
