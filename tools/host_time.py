"""Measures host enqueue time per step (no sync inside) vs device time."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from probabilisticteacher_b200.config import c2f_config
from probabilisticteacher_b200.engine.trainer import PTrainer
dev = torch.device("cuda:0")
cfg = c2f_config(); cfg.UNSUPNET.BURN_UP_STEP = 0
pool = bench.synthetic_pool(2, 2, 800, 1333, 8, 1234, device=dev)
tr = PTrainer(cfg, bench.cycle(pool), device=dev)
for _ in range(3): tr.run_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): tr.run_step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/10:.2f} ms/step; total {1e3*(t2-t0)/10:.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): tr.run_step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
