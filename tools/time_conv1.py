import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticteacher_b200 import ops
from probabilisticteacher_b200._lib import call
dev = torch.device("cuda:0")
N, H, W = 4, 800, 1333
img = torch.randint(0, 256, (N, 3 * H * W), dtype=torch.uint8, device=dev)
hw = torch.tensor([[H, W]] * N, dtype=torch.int32, device=dev)
wp = torch.randn(64, 32, device=dev).half()
b = torch.zeros(64, device=dev)
mean, std = (103.53, 116.28, 123.675), (1.0, 1.0, 1.0)
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("conv1_u8 fused ms (4 img):", t(lambda: ops.conv1_u8(img, hw, H, W, mean, std, wp, b)))
print("preprocess_im2col ms (4 img):", t(lambda: ops.preprocess_im2col(img, hw, H, W, mean, std)))
