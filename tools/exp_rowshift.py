import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess
# the probe is NOT part of libptb200.so: it is compiled here, next to the product's tensor-map helper
_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_csrc = os.path.join(_root, "probabilisticteacher_b200", "csrc")
_so = os.path.join(_root, "tools", "_exp_rowshift.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                       "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-I", _csrc,
                       os.path.join(_root, "tools", "exp_rowshift.cu"), os.path.join(_csrc, "gemm_tn.cu"), "-o", _so])
L = ctypes.CDLL(_so)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
A = torch.randn(136, 64, generator=g).half().to(dev)
B = torch.randn(64, 64, generator=g).half().to(dev)
for ubo in (0, 1):
    out = torch.zeros(8, 128, 64, device=dev)
    rc = L.ptb200_exp_rowshift(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(out.data_ptr()), ubo, ctypes.c_void_p(0))
    torch.cuda.synchronize()
    for s in range(8):
        ref = A[s:s + 128].float() @ B.float().t()
        err = float((out[s] - ref).abs().max() / ref.abs().max())
        print(f"use_base_offset={ubo} shift {s}: rel err {err:.3e}")
