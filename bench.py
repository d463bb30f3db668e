#!/usr/bin/env python
"""Benchmark of the Probabilistic Teacher per-step hot path (BASELINE.json metric:
teacher+student training iters/sec @ 3x800x1333, bs=2 source + 2 target pairs per GPU).

  python bench.py --gpus N --steps K --warmup W          # this repo's B200 path
  python bench.py --impl reference --gpus N ...          # the reference algorithm on host CPU cores

One "step" = one post-burn-in PTrainer.run_step (pt/engine/trainer.py:291-392): EMA teacher update,
teacher forward on 2 weak target images, student supervised forward on 4 source images, student
unsupervised forward on 2 strong target images, backward, gradient all-reduce, clip, SGD.
`value` times the step with the uint8 images already resident in HBM; `e2e` times the same step
through the public trainer API with pinned HOST images (H2D copies inside the timed region) and a
D2H read of the 8 loss scalars every step. `value` = iterations/s summed over ranks (each rank runs
its own bs=2+2 iteration: weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H_IMG, W_IMG = 800, 1333
PAIRS_PER_GPU = 2
METRIC = "teacher+student training iters/sec @ 3x800x1333 bs=2/GPU"


# ------------------------------------------------------------------------------------------ data
def synthetic_pool(n_batches, pairs, H, W, num_classes, seed, device=None, pin=True):
    """SURVEY.md 8d synthetic inputs: uint8 uniform images, 12 GT boxes per source image. Returns a list
    of (label_q, label_k, unlabel_q, unlabel_k) tuples of dict lists in the reference's format."""
    import torch
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    from probabilisticteacher_b200.synthetic import synthetic_batch
    pool = []
    for b in range(n_batches):
        lab = synthetic_batch(pairs, H, W, num_classes, seed + 2 * b)
        unl = synthetic_batch(pairs, H, W, num_classes, seed + 2 * b + 1, labelled=False)

        def conv(batch, with_inst):
            out = []
            for d in batch:
                img = d["image"]
                if device is not None:
                    img = img.to(device)
                elif pin:
                    img = img.pin_memory()
                nd = {"image": img, "height": H, "width": W}
                if with_inst:
                    i = d["instances"]
                    nd["instances"] = FreeInstances((H, W), gt_boxes=Boxes(i.gt_boxes.tensor.clone()),
                                                    gt_classes=i.gt_classes.clone())
                out.append(nd)
            return out
        lq = conv(lab, True)
        lk = conv(lab, True)
        uq = conv(unl, False)
        uk = conv(unl, False)
        pool.append((lq, lk, uq, uk))
    return pool


def cycle(pool):
    i = 0
    while True:
        lq, lk, uq, uk = pool[i % len(pool)]
        yield ([dict(d) for d in lq], [dict(d) for d in lk], [dict(d) for d in uq], [dict(d) for d in uk])
        i += 1


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nme in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------ roofline profiler
class GemmProfiler:
    """Times every launch of the tcgen05 GEMM kernels with CUDA events on the launching stream and
    books the algorithmic FLOPs (2*MACs)."""

    def __init__(self):
        self.records = []
        self.shapes = []
        self.other = []

    def dump(self, path):
        rows = []
        for (name, flops, e0, e1), sh in zip(self.records, self.shapes):
            ms = e0.elapsed_time(e1)
            rows.append({"kind": sh[0], "shape": sh[1:], "ms": ms, "tflops": flops / ms / 1e9})
        os.makedirs(os.path.dirname(path), exist_ok=True)
        json.dump(rows, open(path, "w"))

    def begin(self, name, args):
        import torch
        if name == "ptb200_gemm_tn_f16":
            batch, rows, k, taps, n_total = args[1], args[2], args[3], args[6], args[9]
            n_valid = args[25] if args[11] in (2, 4) else n_total
            flops = 2.0 * batch * rows * k * taps * n_valid
        elif name == "ptb200_gemm_wgrad_f16":
            batch, rows, m, n, taps = args[6], args[7], args[8], args[9], args[10]
            flops = 2.0 * batch * rows * m * n * taps
        else:
            # every other entry point: time only (reported as kernels_ms_per_step); ROIAlign also books its
            # algorithmic bytes = live rois x 7 x 7 x C x 2 B (the K-major fc1 operand written / read once)
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            extra = None
            if name in ("ptb200_roi_align_fwd_f16", "ptb200_roi_align_bwd_f16"):
                extra = (args[6], args[7], args[9] * args[9] * args[4] * 2)  # counts tensor, cap, bytes per roi
            self.other.append((name, e0, e1, extra))
            return ("other", None, e0, e1)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        if name == "ptb200_gemm_tn_f16":
            self.shapes.append(("tn", args[1], args[2], args[3], args[6], args[9], args[10], args[11]))
        else:
            self.shapes.append(("wgrad", args[6], args[7], args[8], args[9], args[10], 0, 0))
        return (name, flops, e0, e1)

    def end(self, tok):
        if tok is not None:
            tok[3].record()
            if tok[0] != "other":
                self.records.append(tok)

    def other_summary(self):
        """({entry point: ms per step}, ROIAlign {name: (ms, GB/s)})."""
        ms = {}
        roi = {}
        for name, e0, e1, extra in self.other:
            t = e0.elapsed_time(e1)
            ms[name] = ms.get(name, 0.0) + t
            if extra is not None:
                counts, cap, per_roi = extra
                live = int(counts.clamp(max=cap).sum()) if counts is not None else 0
                a = roi.setdefault(name, [0.0, 0.0])
                a[0] += t
                a[1] += live * per_roi
        return ms, {k: {"ms_per_step": v[0], "algorithmic_GBps": v[1] / v[0] / 1e6 if v[0] > 0 else 0.0,
                        "MB_per_step": v[1] / 1e6} for k, v in roi.items()}

    def summary(self):
        tot_f = {"ptb200_gemm_tn_f16": 0.0, "ptb200_gemm_wgrad_f16": 0.0}
        tot_t = {"ptb200_gemm_tn_f16": 0.0, "ptb200_gemm_wgrad_f16": 0.0}
        cnt = {"ptb200_gemm_tn_f16": 0, "ptb200_gemm_wgrad_f16": 0}
        for name, flops, e0, e1 in self.records:
            tot_f[name] += flops
            tot_t[name] += e0.elapsed_time(e1) * 1e-3
            cnt[name] += 1
        return tot_f, tot_t, cnt


# ------------------------------------------------------------------------------------------ CPU oracle leg
def cpu_oracle_iters_per_s(steps, warmup, pairs=1, H=H_IMG, W=W_IMG):
    """Times the CPU restatement of the reference step (oracle/pt_oracle.py, all host threads) on a
    bounded sample: `pairs` source + `pairs` target images per step. Returns (iters/s normalised to a
    bs=2+2 iteration, cores, sample description)."""
    import torch
    from oracle import pt_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.OracleCfg()
    student = O.OracleRCNN(cfg, seed=1)
    teacher = O.OracleRCNN(cfg, seed=1)
    opt = O.make_optimizer(student, cfg)
    times = []
    for s in range(warmup + steps):
        lq = O.synthetic_batch(pairs, H, W, cfg.num_classes, 1234 + 2 * s)
        lk = [dict(d) for d in lq]
        uq = O.synthetic_batch(pairs, H, W, cfg.num_classes, 1235 + 2 * s, labelled=False)
        uk = [dict(d) for d in uq]
        t0 = time.perf_counter()
        O.run_step(student, teacher, opt, (lq, lk, uq, uk), cfg, [0.75] * pairs, [0.75] * pairs)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    iters_per_s = (pairs / PAIRS_PER_GPU) / per_step
    sample = (f"{len(times)} oracle step(s) of {pairs} source + {pairs} target 3x{H}x{W} images "
              f"({per_step:.2f} s/step), scaled to a {PAIRS_PER_GPU}+{PAIRS_PER_GPU} iteration")
    return iters_per_s, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 2))
    warm = min(args.warmup, 1)
    v, cores, sample = cpu_oracle_iters_per_s(steps, warm, H=args.height, W=args.width)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": 1000.0 / v if v > 0 else None, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CitysScape2FoggyCityscape config, synthetic 3x{args.height}x{args.width}, 2 source + "
                               "2 target per iteration (CPU run on a 1+1 sample, scaled)", "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": v, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "detectron2 is not installable here (no package, no network): the reference arm is the CPU "
                "oracle port of the reference path (oracle/pt_oracle.py), all host threads; steps capped at "
                f"{steps} timed + {warm} warm-up to stay within minutes",
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ main arm
def timed_steps(trainer, steps, dist, device, read_losses=False, host_sink=None):
    import torch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses = trainer.step()
        if read_losses:
            vec = torch.stack([losses[k].reshape(()) for k in sorted(losses)])
            host_sink.copy_(vec, non_blocking=False)
    e1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ptb200", choices=["ptb200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true")
    ap.add_argument("--no-concurrent", action="store_true")
    ap.add_argument("--config", default="c2f", choices=["c2f", "k2c"],
                    help="c2f: BASELINE configs 2/3 (default, the headline); k2c: config 4 (K = 1; use with "
                         "--height 600 --width 2000)")
    ap.add_argument("--height", type=int, default=H_IMG)
    ap.add_argument("--width", type=int, default=W_IMG)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as tdist
    from probabilisticteacher_b200 import _lib
    from probabilisticteacher_b200.config import c2f_config, k2c_config
    from probabilisticteacher_b200.engine.trainer import PTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        tdist.init_process_group("nccl", device_id=device)
        dist = tdist
    warmup = max(args.warmup, 3)
    cfg = k2c_config() if args.config == "k2c" else c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0  # time the post-burn-in (teacher + student) iteration
    H, W = args.height, args.width
    K = cfg.MODEL.ROI_HEADS.NUM_CLASSES

    pool_dev = synthetic_pool(2, PAIRS_PER_GPU, H, W, K, 1234 + 100 * rank, device=device)
    pool_host = synthetic_pool(2, PAIRS_PER_GPU, H, W, K, 1234 + 100 * rank, device=None, pin=True)
    use_graph = not args.no_cuda_graph
    trainer = PTrainer(cfg, cycle(pool_dev), device=device, seed=0, use_cuda_graph=use_graph,
                       concurrent=use_graph and not args.no_concurrent)

    for _ in range(warmup + (5 if use_graph else 0)):
        trainer.step()
    torch.cuda.synchronize()

    # ---- device-resident arm
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count[0]
    ms = timed_steps(trainer, args.steps, dist, device)
    launches = (_lib.launch_count[0] - l0) // args.steps  # eager launches; graph replays are counted below
    ms_per_step = ms / args.steps
    value = world * 1000.0 / ms_per_step

    # ---- end-to-end arm: pinned host images in, loss scalars out, every step
    trainer._data_loader_iter = cycle(pool_host)
    host_sink = torch.empty(8, dtype=torch.float32).pin_memory()
    trainer.step()
    ms_e2e = timed_steps(trainer, args.steps, dist, device, read_losses=True, host_sink=host_sink)
    e2e_value = world * 1000.0 * args.steps / ms_e2e
    clk = clocks.stop() if rank == 0 else None  # sampled over BOTH timed regions (device-resident and end-to-end)
    h2d = 4 * PAIRS_PER_GPU * 3 * H * W  # label_q, label_k, unlabel_q, unlabel_k uint8 images
    d2h = 8 * 4

    # ---- roofline of the dominant kernel (tcgen05 implicit GEMM), one extra profiled step
    prof = GemmProfiler()
    _lib.profiler[0] = prof
    trainer._data_loader_iter = cycle(pool_dev)
    l1 = _lib.launch_count[0]
    trainer.run_step()  # eager: every kernel is launched (and timed) individually
    eager_launches = _lib.launch_count[0] - l1
    torch.cuda.synchronize()
    _lib.profiler[0] = None
    tot_f, tot_t, cnt = prof.summary()
    other_ms, roi_stats = prof.other_summary()
    if os.environ.get("PTB_DUMP_GEMM") and rank == 0:
        prof.dump(os.path.join(ROOT, "gpurun_out", "gemm_launches.json"))
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained"
    gt = tot_t["ptb200_gemm_tn_f16"]
    achieved = tot_f["ptb200_gemm_tn_f16"] / gt / 1e12 if gt > 0 else 0.0
    wt = tot_t["ptb200_gemm_wgrad_f16"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r1_gemm_dram_traffic.json")
    if os.path.exists(tp):  # committed summary of an ncu pass over one step (dram bytes per launch)
        tj = json.load(open(tp)).get("void gemm_tn_kernel<0>")
        if tj:
            traffic = (tj["dram_read_MB_per_launch"] + tj["dram_write_MB_per_launch"]) * 1e6
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_note": "mean DRAM read+write bytes per gemm_tn_kernel launch, ncu pass over one step "
                                "(profiles/r1_gemm_dram_traffic.json)",
                "kernel": "gemm_tn_kernel",
                "launches_per_step": cnt["ptb200_gemm_tn_f16"], "kernel_ms_per_step": gt * 1e3,
                "peak_source": peak_src,
                "wgrad_kernel": {"achieved": tot_f["ptb200_gemm_wgrad_f16"] / wt / 1e12 if wt > 0 else 0.0,
                                 "kernel_ms_per_step": wt * 1e3, "launches_per_step": cnt["ptb200_gemm_wgrad_f16"]},
                "share_of_step": (gt + wt) * 1e3 / ms_per_step}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_oracle_iters_per_s(1, 0)
        cpu_baseline = {"value": v, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "iters/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16",
            "dtype_detail": "fp16 operands / fp32 accumulate (tcgen05 kind::f16), fp32 master weights, losses and "
                            "optimizer; the reference runs fp32 with AMP off",
            "data": "synthetic",
            "config": {"workload": ("KITTI2CitysScape config (configs/pt/final_k2c.yaml, K = 1), " if args.config == "k2c" else
                                    "CitysScape2FoggyCityscape config (configs/pt/final_c2f.yaml + train.sh overrides), ") +
                                   f"synthetic 3x{H}x{W}, {PAIRS_PER_GPU} source + {PAIRS_PER_GPU} target pairs per GPU, "
                                   "full post-burn-in PT iteration",
                       "global_batch": f"{PAIRS_PER_GPU * world}+{PAIRS_PER_GPU * world}",
                       "parallelism": f"dp{world}",
                       "l2": "per-step working set (GBs of activations) is far larger than the 126 MB L2",
                       "value_definition": "iterations/s summed over ranks (each rank runs one 2+2 iteration per step)"},
            "pairs_per_s": value * PAIRS_PER_GPU,
            "e2e": {"value": e2e_value, "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": eager_launches, "cuda_graph": use_graph, "concurrent_branches": use_graph and not args.no_concurrent,
            "gpu_launches_note": "kernels of libptb200.so per step (counted on an eager step; in CUDA-graph mode the "
                                 "same kernels are replayed from the captured graph)",
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernels_ms_per_step": {k.replace("ptb200_", ""): round(v, 4) for k, v in
                                    sorted(other_ms.items(), key=lambda kv: -kv[1])[:12]},
            "roialign": roi_stats,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
