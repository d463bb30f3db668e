"""BASELINE config 5: backbone-only micro-benchmark -- VGG16 13-conv stack forward (fused pre-processing +
conv1_1, 12 tcgen05 implicit-GEMM convs, 4 max-pools) on N x 3 x 1024 x 2048 uint8 images, fp16 operands.
Prints one JSON line: ms, TFLOP/s (1282.85 GFLOP per image) and the fraction of the sustained /
burst bf16 peaks. Run on the GPU box:  python tools/bench_backbone.py [N] [H] [W]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probabilisticteacher_b200.config import c2f_config  # noqa: E402
from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model  # noqa: E402


def conv_flops(H, W):
    chans = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 512), (512, 512),
             (512, 512), (512, 512), (512, 512), (512, 512)]
    pool_after = {1, 3, 6, 9}
    f = 0.0
    h, w = H, W
    for i, (ci, co) in enumerate(chans):
        f += 2.0 * h * w * ci * co * 9
        if i in pool_after:
            h, w = h // 2, w // 2
    return f


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    dev = torch.device("cuda:0")
    model = build_model(c2f_config(), dev, with_grads=False)
    model.init_synthetic(0)
    model.train()
    g = torch.Generator().manual_seed(1)
    batch = [{"image": torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8).to(dev)} for _ in range(N)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run():
        act, _, _ = model.preprocess_image(batch)
        return model.backbone(act, save=False)[0]["vgg_block5"]

    with torch.no_grad():
        for _ in range(3):
            run()
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    fl = conv_flops(H, W) * N
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    sus, burst = peaks.get("bf16_tflops_sustained", 1400.0), peaks.get("bf16_tflops_burst", 1650.0)
    tf = fl / ms / 1e9
    print(json.dumps({"workload": f"VGG16 conv stack fwd, {N} x 3x{H}x{W}, fp16 operands", "ms": ms,
                      "gflop": fl / 1e9, "tflops": tf, "frac_sustained": tf / sus, "frac_burst": tf / burst,
                      "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 1400 / 1650 TFLOP/s",
                      "l2": "256 MB flush between timed runs"}))


if __name__ == "__main__":
    main()
