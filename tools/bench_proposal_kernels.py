"""Stand-alone timing of the non-GEMM kernels DESIGN.md section 7 ranks first (ROIAlign forward / backward, the
proposal selection chain = key build + radix sort + decode + sort + NMS + gather, and NMS alone) at the shapes of
one bench step (3x800x1333: feature map 50x83, 37 350 anchors, 12 000 pre-NMS / 2 000 post-NMS candidates per
image), without building the models: a kernel iteration costs seconds of GPU time instead of a full bench run.
CUDA events on the launching stream, L2 flushed between timed launches, median of 7. Run on the GPU box:

    python tools/bench_proposal_kernels.py            # all
    ncu --set full -k regex:roi_align_roi_kernel -s 3 -c 1 python tools/bench_proposal_kernels.py roialign

Proposal boxes are anchors perturbed like the RPN output at the synthetic initialisation (deltas ~ N(0, 0.2)),
i.e. mostly 128-512 pixel boxes: the regime the step runs in."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from probabilisticteacher_b200 import ops  # noqa: E402
from probabilisticteacher_b200.config import c2f_config  # noqa: E402
from probabilisticteacher_b200.modeling.proposal_generator.proposal_utils import find_top_rpn_proposals  # noqa: E402

DEV = torch.device("cuda:0")
H, W, A, C = 50, 83, 9, 512
IMG_H, IMG_W = 800.0, 1333.0
_flush = None


def timed(fn, reps=7):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3  # us


def anchors_cpu():
    wh = torch.tensor(c2f_config().MODEL.ANCHOR_GENERATOR.ANCHOR[0])
    cell = torch.stack([-wh[:, 0] / 2, -wh[:, 1] / 2, wh[:, 0] / 2, wh[:, 1] / 2], -1)
    ys, xs = torch.meshgrid(torch.arange(H) * 16.0, torch.arange(W) * 16.0, indexing="ij")
    shifts = torch.stack([xs, ys, xs, ys], -1).reshape(-1, 1, 4)
    return (shifts + cell.view(1, -1, 4)).reshape(-1, 4)


def rpn_like_inputs(N, g):
    lg = torch.zeros(N, H, W + 1, A)
    lg[:, :, :W] = torch.randn(N, H, W, A, generator=g) * 0.05
    dl = torch.zeros(N, H, W + 1, A * 8)
    dl[:, :, :W] = torch.randn(N, H, W, A * 8, generator=g) * 0.2
    return lg.reshape(N, -1, A).to(DEV), dl.reshape(N, -1, A * 8).to(DEV)


def bench_proposals(g):
    anchors = anchors_cpu().to(DEV)
    flag = torch.zeros(1, dtype=torch.int32, device=DEV)
    for N in (2, 4):
        lg, dl = rpn_like_inputs(N, g)
        hw = torch.tensor([[IMG_H, IMG_W]] * N, device=DEV)
        us = timed(lambda: find_top_rpn_proposals(lg, dl, anchors, N, H, W, A, hw, 0.7, 12000, 2000, 0.0, flag))
        boxes, scores, count = find_top_rpn_proposals(lg, dl, anchors, N, H, W, A, hw, 0.7, 12000, 2000, 0.0, flag)
        print(f"find_top_rpn_proposals N={N}: {us:8.1f} us   kept {count.tolist()}")
    return boxes, count  # N = 4 proposals, reused as rois


def bench_nms(g):
    for N, n in ((2, 12000), (4, 12000), (2, 16000)):
        centers = torch.rand(max(n // 20, 4), 2, generator=g) * torch.tensor([IMG_W - 300, IMG_H - 300])
        c = centers[torch.randint(0, centers.shape[0], (N, n), generator=g)]
        wh = torch.rand(N, n, 2, generator=g) * 300 + 30
        b = torch.cat([c, c + wh], -1) + torch.randn(N, n, 4, generator=g) * 6
        cap = (n + 63) // 64 * 64
        order = torch.zeros(N, cap, dtype=torch.int32)
        order[:, :n] = torch.arange(n, dtype=torch.int32)
        bd, od = b.to(DEV).contiguous(), order.to(DEV)
        cnt = torch.full((N,), n, dtype=torch.int32, device=DEV)
        us = timed(lambda: ops.nms(bd, od, cnt, 0.7, 2000))
        _, kc = ops.nms(bd, od, cnt, 0.7, 2000)
        print(f"nms N={N} n={n}: {us:8.1f} us   kept {kc.tolist()}")


def bench_roialign(g, rois=None, count=None):
    for name, N, cap in (("teacher 2 x 2000", 2, 2000), ("supervised 4 x 512", 4, 512)):
        feat = ops.FlatAct((torch.randn(N, H * (W + 1), C, generator=g) * 0.5).half().to(DEV), H, W)
        if rois is not None and rois.shape[0] >= N and rois.shape[1] >= cap:
            r = rois[:N, :cap].contiguous()
            cnt = count[:N].clamp(max=cap).to(torch.int32)
        else:  # anchors + noise, clipped to the image
            a = anchors_cpu()
            pick = torch.randint(0, a.shape[0], (N, cap), generator=g)
            r = a[pick] + torch.randn(N, cap, 4, generator=g) * 8
            r[..., 0::2] = r[..., 0::2].clamp(0, IMG_W)
            r[..., 1::2] = r[..., 1::2].clamp(0, IMG_H)
            r = r.to(DEV).contiguous()
            cnt = torch.full((N,), cap, dtype=torch.int32, device=DEV)
        live = int(cnt.sum())
        mb = live * 49 * C * 2 / 1e6
        us = timed(lambda: ops.roi_align_fwd(feat, r, cnt, cap, 1.0 / 16, 7))
        print(f"roi_align fwd {name}: {us:8.1f} us   {live} live rois, {mb:.0f} MB out -> {mb / us * 1e3:.0f} GB/s algorithmic")
        dout = (torch.randn(N * cap, 49 * C, generator=g) * 0.01).half().to(DEV)
        us = timed(lambda: ops.roi_align_bwd(dout, feat, r, cnt, cap, 1.0 / 16, 7))
        print(f"roi_align bwd {name}: {us:8.1f} us   (includes zero-filling the {N * H * (W + 1) * C * 4 / 1e6:.0f} MB fp32 gradient map)"
              f" -> {mb / us * 1e3:.0f} GB/s algorithmic")


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    g = torch.Generator().manual_seed(0)
    rois = count = None
    if which in ("all", "proposals"):
        rois, count = bench_proposals(g)
    if which in ("all", "nms"):
        bench_nms(g)
    if which in ("all", "roialign"):
        bench_roialign(g, rois, count)


if __name__ == "__main__":
    main()
