"""Diagnostics of the f16x3 parity precision: accumulation bias of the tensor core vs K-chunking, and the
end-to-end report of tests/test_parity_x3_gpu.py at a chosen size."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from probabilisticteacher_b200 import ops  # noqa: E402

cuda = torch.device("cuda:0")


def accumulation_bias():
    g = torch.Generator().manual_seed(0)
    rows, N = 512, 128
    for K in (512, 4608, 25088):
        A = (torch.rand(rows, K, generator=g) * 4).to(cuda)
        Wt = (torch.rand(N, K, generator=g) * 0.05).to(cuda)
        A3 = ops.split3_pack(A, K, 1.0, 0).view(1, rows, 3 * K)
        W3 = ops.split3_pack(Wt, K, 4096.0, 1)
        ref = A.double() @ Wt.double().t()
        for chunk in (0, 8, 4, 2, 1):
            ops.X3_MAX_K_ITERS[0] = max(chunk, 1)
            o = ops.split3_unpack(ops.gemm_tn_x3(A3, W3, 1.0 / 4096.0, epi=ops.EPI_SPLIT3,
                                                 ksplit=1 if chunk == 0 else None), N).double()
            e = (o - ref) / ref
            print(f"K={K:6d} k-iters per chunk={chunk or 'all'}: mean signed rel err {float(e.mean()):+.3e}  max |rel| {float(e.abs().max()):.3e}")
        o16 = ops.gemm_tn(A.half().view(1, rows, K), Wt.half(), epi=ops.EPI_BIAS).double().view(rows, N)
        e = (o16 - ref) / ref
        print(f"K={K:6d} plain f16: mean signed rel err {float(e.mean()):+.3e}  max |rel| {float(e.abs().max()):.3e}")


if __name__ == "__main__":
    accumulation_bias()
    ops.X3_MAX_K_ITERS[0] = int(os.environ.get("X3_CHUNK", "2"))
    import test_parity_x3_gpu as T
    for args in ((192, 272, 8, "DifferentiableAnchorGenerator", 2, 3), (144, 240, 1, "DefaultAnchorGenerator", 2, 5),
                 (800, 1333, 8, "DefaultAnchorGenerator", 1, 3)):
        if len(sys.argv) > 1 and sys.argv[1] == "small" and args[0] == 800:
            continue
        rep = T._full_iteration(cuda, *args)
        print(args)
        for k, v in rep.items():
            if k == "teacher":
                for d in v:
                    print("   teacher", json.dumps(d))
            else:
                print(f"   {k:32s} gpu {v[0]:.7f} oracle {v[1]:.7f} rel {abs(v[0]-v[1])/max(abs(v[1]), 1e-12):.2e}")
