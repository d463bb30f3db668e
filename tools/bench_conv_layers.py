"""Micro-benchmark of the implicit-GEMM kernel on every VGG16 / RPN / box-head contraction at
3x800x1333 (CUDA events, L2 flushed between timed launches). Run on the GPU box."""
import ctypes
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticteacher_b200._lib import lib, ptr, stream_ptr, check


def run(A, B, bias, D, taps, shifts, bn, W, Wp, epi=0):
    batch, rows, lda = A.shape
    n_total = B.shape[0]
    k = B.shape[1] // taps
    sh = (ctypes.c_int * 9)(*(list(shifts) + [0] * (9 - len(shifts))))
    rc = lib().ptb200_gemm_tn_f16(
        ptr(A), batch, rows, k, ctypes.c_int64(lda), ctypes.c_int64(rows * lda), taps, sh, ptr(B),
        n_total, bn, epi, ptr(bias), bias.numel(), ptr(D), ctypes.c_int64(n_total),
        ctypes.c_int64(rows * n_total), None, W, Wp, None, 0, None, 0, 0, 0, 0, 1, None, 0, stream_ptr())
    check(rc)


def main():
    dev = torch.device("cuda:0")
    N = int(os.environ.get("NIMG", "2"))
    H0, W0 = 800, 1333
    layers = [("1_2", 64, 64, 1), ("2_1", 64, 128, 2), ("2_2", 128, 128, 2), ("3_1", 128, 256, 4),
              ("3_2", 256, 256, 4), ("4_1", 256, 512, 8), ("4_2", 512, 512, 8), ("5_1", 512, 512, 16)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = []
    only = os.environ.get("LAYER")
    for name, cin, cout, div in layers:
        if only and name != only:
            continue
        H, W = H0 // div, W0 // div
        Wp = W + 1
        A = torch.randn(N, H * Wp, cin, device=dev).half()
        B = (torch.randn(cout, 9 * cin, device=dev) / (3 * cin ** 0.5)).half()
        bias = torch.zeros(cout, device=dev)
        D = torch.empty(N, H * Wp, cout, device=dev, dtype=torch.float16)
        shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
        bn = min(cout, 256)
        for _ in range(3):
            run(A, B, bias, D, 9, shifts, bn, W, Wp)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            run(A, B, bias, D, 9, shifts, bn, W, Wp)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        fl = 2.0 * N * H * W * cin * cout * 9
        res.append(dict(layer=name, H=H, W=W, cin=cin, cout=cout, ms=ms, tflops=fl / ms / 1e9))
        print(res[-1], flush=True)
    # box head fc1: 4000 x 25088 x 1024
    if only and not only.startswith("fc"):
        return
    for name, M, K, Nn in [("fc1_teacher", 4000, 25088, 1024), ("fc1_sup", 2048, 25088, 1024), ("fc2", 4000, 1024, 1024)]:
        if only and name != only:
            continue
        A = torch.randn(1, M, K, device=dev).half()
        B = (torch.randn(Nn, K, device=dev) / K ** 0.5).half()
        bias = torch.zeros(Nn, device=dev)
        D = torch.empty(1, M, Nn, device=dev, dtype=torch.float16)
        for _ in range(3):
            run(A, B, bias, D, 1, [0], 256, 0, 0)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            run(A, B, bias, D, 1, [0], 256, 0, 0)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        res.append(dict(layer=name, ms=ms, tflops=2.0 * M * K * Nn / ms / 1e9))
        print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_conv_layers.json", "w"), indent=1)


if __name__ == "__main__":
    main()
