"""Dev (GPU): how much of the 12 000-row candidate list does the RPN NMS scan visit in the benchmark's regime (randomly
initialised detector at 800x1333)? Reports survivors per image, the row at which max_keep is reached, and the share of
64-row blocks without any survivor (those could be skipped without touching their mask rows)."""
import os
import sys

import torch
import torchvision

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from probabilisticteacher_b200 import ops  # noqa: E402
from probabilisticteacher_b200._lib import call  # noqa: E402
from probabilisticteacher_b200.config import c2f_config  # noqa: E402
from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model  # noqa: E402
from probabilisticteacher_b200.synthetic import synthetic_batch  # noqa: E402

dev = torch.device("cuda:0")
H, W = 800, 1333
model = build_model(c2f_config(), dev, precision="f16", with_grads=False)
model.init_synthetic(0)
model.train()
batch = synthetic_batch(2, H, W, 8, 5, labelled=False)
with torch.no_grad():
    act, sizes, img_hw = model.preprocess_image(batch)
    feats, _ = model.backbone(act, save=False)
    feat = feats["vgg_block5"]
    rpn = model.proposal_generator
    anchors = rpn.anchor_generator(feat.H, feat.W)
    t, logits, deltas = rpn.rpn_head(feat)
    N, A = 2, 9
    R = feat.H * feat.W * A
    k = 12000
    keys = torch.empty(N, R, dtype=torch.int32, device=dev)
    vals = torch.empty(N, R, dtype=torch.int32, device=dev)
    call("ptb200_rpn_make_keys", logits, logits.shape[2], N, feat.H, feat.W, A, keys, vals)
    ops.segmented_sort(keys, vals)
    boxes = torch.empty(N, k, 4, device=dev)
    scores = torch.empty(N, k, device=dev)
    keys2 = torch.empty(N, k, dtype=torch.int32, device=dev)
    vals2 = torch.empty(N, k, dtype=torch.int32, device=dev)
    valid = torch.empty(N, dtype=torch.int32, device=dev)
    call("ptb200_rpn_topk_decode", vals, R, logits, logits.shape[2], deltas, deltas.shape[2], anchors, N, feat.H, feat.W,
         A, k, img_hw, 0.0, boxes, scores, keys2, vals2, valid, rpn.nonfinite_flag)
    ops.segmented_sort(keys2, vals2)
    for n in range(N):
        c = int(valid[n])
        order = vals2[n, :c].long()
        b = boxes[n][order]
        s = scores[n][order]
        keep = torchvision.ops.nms(b, s, 0.7)   # b is already in descending-score order: keep = positions
        keep, _ = keep.sort()
        kept = keep.numel()
        stop_row = int(keep[1999]) if kept >= 2000 else c
        blocks = (stop_row + 63) // 64
        per_block = torch.bincount((keep[keep < stop_row] // 64), minlength=blocks)
        print(f"image {n}: {c} candidates, {kept} survive NMS(0.7); the scan reaches 2000 survivors at row {stop_row} "
              f"({blocks} blocks); blocks without a survivor: {int((per_block == 0).sum())}, mean survivors per block "
              f"{float(per_block.float().mean()):.1f}")
