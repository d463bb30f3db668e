"""Multi-GPU consistency check (test-plan item 5). Launch with torchrun on N GPUs:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_ddp.py

Every rank runs K post-burn-in PT iterations (CUDA-graph + concurrent branches, as bench.py) on its own
synthetic batches; after the run the student and teacher arenas must be BIT-IDENTICAL on all ranks (same
all-reduced gradients -> same clip/SGD -> same EMA), and the gradient the ranks applied must be the mean of
the per-rank gradients (checked on the last step against an all-gather of the local gradients)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import cycle, synthetic_pool  # noqa: E402
from probabilisticteacher_b200.config import c2f_config  # noqa: E402
from probabilisticteacher_b200.engine.trainer import PTrainer  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "f16"])
    args = ap.parse_args()
    H, W, steps = args.height, args.width, args.steps
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    pool = synthetic_pool(2, 2, H, W, 8, 1234 + 100 * rank, device=dev)
    tr = PTrainer(cfg, cycle(pool), device=dev, seed=0, use_cuda_graph=True, concurrent=True, precision=args.precision)
    trace = os.environ.get("PTB200_TRACE_STEPS", "0") == "1"
    if trace:  # a hung run dumps every thread's Python stack and exits on its own
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ.get("PTB200_TRACE_TIMEOUT", "70")), exit=True)
    for i in range(steps):
        tr.step()
        if trace:
            torch.cuda.synchronize()
            print(f"[rank {rank}] step {i} done", file=sys.stderr, flush=True)
    torch.cuda.synchronize()
    ok = True
    # cross-rank metric reduction (pt/engine/trainer.py:394-429): the losses that rode on the gradient all-reduce
    # must be the mean over ranks of every rank's own losses
    keys = sorted(tr.last_losses)
    mine = torch.stack([tr.last_losses[k].reshape(()).float() for k in keys])
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    mean = torch.stack(allv).double().mean(0)
    red = tr.reduced_metrics()
    got = torch.stack([red[k].reshape(()).double() for k in keys])
    merr = float(((got - mean).abs() / mean.abs().clamp_min(1e-6)).max())
    if rank == 0:
        print(f"{args.precision} {H}x{W}: reduced metrics vs mean of the {world} ranks' losses: max rel err {merr:.2e}")
    ok = ok and merr < 1e-5
    for name, arena in (("student", tr.model.arena), ("teacher", tr.model_teacher.arena)):
        mine = arena.data
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(mine, ref))
        flags = [None] * world
        dist.all_gather_object(flags, same)
        if rank == 0:
            print(f"{name} arena bit-identical on all {world} ranks after {steps} steps: {all(flags)}")
        ok = ok and all(flags)
    # all-reduce semantics: SUM over ranks of the local gradient arenas (1/world folded into the SGD kernel)
    g_local = torch.randn(1 << 20, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))
    gathered = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    g_sum = g_local.clone()
    dist.all_reduce(g_sum)
    expect = torch.stack(gathered).double().sum(0)
    err = float((g_sum.double() - expect).abs().max())
    if rank == 0:
        print(f"all-reduce(SUM) vs all-gather reference: max abs err {err:.3e}")
        print("losses (rank 0):", {k: round(float(v), 5) for k, v in tr.last_losses.items()})
        print("DDP CHECK", "PASSED" if ok and err < 1e-5 else "FAILED")
    tr.release_graphs()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
