"""CPU oracle for the Probabilistic Teacher hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch (CPU, fp32) restatement of the reference algorithm for the per-step hot path. Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module; the product package never does.

Every function cites the reference file:line it follows (paths relative to the reference checkout
of hikvision-research/ProbabilisticTeacher). The arithmetic the reference inherits from
detectron2==0.5 / torchvision (un-vendored third-party dependencies, absent from the reference
tree) is restated from those releases' published behaviour and marked "d2 v0.5".

Pinning: `oracle/make_golden.py` imports the reference's own pt/modeling files (unmodified, from
/root/reference, through the stub package in oracle/d2shim) and freezes their outputs under
tests/golden/; tests/test_oracle_golden.py checks this restatement against those fixtures.
`oracle/make_golden_model.py` additionally constructs the reference's own model classes
(oracle/d2shim_model.py supplies working detectron2 v0.5 base classes) and runs
GuassianGeneralizedRCNN.forward end to end; tests/test_oracle_golden_model.py checks OracleRCNN
against that (losses bit-identical). `oracle/make_golden_step.py` executes the reference's own
PTrainer.run_step (pt/engine/trainer.py imported unmodified) for two iterations;
tests/test_oracle_golden_step.py checks `run_step` below against it (losses and parameters bit-identical).
`oracle/make_golden_burnin.py` does the same for two source-only (burn-in) iterations (`run_step_burn_in`), and
`oracle/make_golden_eval.py` runs the reference's model classes in eval mode (`forward(..., training=False)` here).
The detectron2 base-class behaviour itself has no golden vectors in the reference (it ships no
tests): that part is pinned only against torchvision ops and analytic cases ("parity unpinned" at
the detectron2 boundary).
"""
import math
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_SCALE_CLAMP = math.log(1000.0 / 16)  # pt/modeling/box_regression.py:28


# --------------------------------------------------------------------------------------------
# containers (duck-typed: .tensor for boxes; has()/attribute access + image_size for instances)
# --------------------------------------------------------------------------------------------
class OBoxes:
    def __init__(self, tensor):
        self.tensor = tensor.reshape(-1, 4).float()

    def __len__(self):
        return self.tensor.shape[0]


class OInst:
    """Minimal stand-in for pt/structures/instances.py:22-46 FreeInstances (no length check)."""

    def __init__(self, image_size, **kw):
        self.image_size = tuple(image_size)
        self._fields = dict(kw)

    def has(self, k):
        return k in self._fields

    def set(self, k, v):
        self._fields[k] = v

    def get(self, k):
        return self._fields[k]

    def __getattr__(self, k):
        if k.startswith("_") or k == "image_size":
            raise AttributeError(k)
        try:
            return self._fields[k]
        except KeyError:
            raise AttributeError(k)

    def __len__(self):
        for v in self._fields.values():
            return len(v)
        return 0


def _bt(b):
    return b.tensor if hasattr(b, "tensor") else b


# --------------------------------------------------------------------------------------------
# hyper-parameters: detectron2 v0.5 defaults + pt/config.py:29-92 + configs/Guassian-RCNN-VGG.yaml
# --------------------------------------------------------------------------------------------
class OracleCfg:
    def __init__(self, **kw):
        self.num_classes = 8                      # configs/Guassian-RCNN-VGG.yaml:24
        self.pixel_mean = (103.530, 116.280, 123.675)   # d2 v0.5 MODEL.PIXEL_MEAN
        self.pixel_std = (1.0, 1.0, 1.0)
        self.anchor_generator = "DifferentiableAnchorGenerator"   # train.sh:9
        self.anchor_sizes = (128.0, 256.0, 512.0)          # configs/Guassian-RCNN-VGG.yaml:11
        self.anchor_ratios = (0.5, 1.0, 2.0)               # configs/Guassian-RCNN-VGG.yaml:12
        self.anchor_wh = [[181.0193, 90.5097], [128.0, 128.0], [90.5097, 181.0193],
                          [362.0387, 181.0193], [256.0, 256.0], [181.0193, 362.0387],
                          [724.0773, 362.0387], [512.0, 512.0], [362.0387, 724.0773]]  # pt/config.py:84-92
        self.anchor_offset = 0.0
        self.stride = 16
        self.rpn_iou_thresholds = (0.3, 0.7)
        self.rpn_iou_labels = (0, -1, 1)
        self.rpn_batch_per_image = 256
        self.rpn_positive_fraction = 0.25         # configs/Guassian-RCNN-VGG.yaml:16
        self.rpn_bbox_weights = (1.0, 1.0, 1.0, 1.0)
        self.rpn_pre_nms_topk = (12000, 6000)     # (train, test)
        self.rpn_post_nms_topk = (2000, 1000)
        self.rpn_nms_thresh = 0.7
        self.rpn_min_size = 0.0
        self.roi_batch_per_image = 512
        self.roi_positive_fraction = 0.25
        self.roi_iou_threshold = 0.5
        self.roi_bbox_weights = (10.0, 10.0, 5.0, 5.0)
        self.roi_score_thresh_test = 0.05
        self.roi_nms_thresh_test = 0.5
        self.detections_per_image = 100
        self.pooler_resolution = 7
        self.fc_dim = 1024
        self.efl = True                            # train.sh:10
        self.efl_lambda = (0.5, 0.5)               # train.sh:11
        self.tau = (0.5, 0.5)                      # train.sh:12
        self.ema_keep_rate = 0.9996                # configs/pt/final_c2f.yaml:27
        self.source_loss_weight = 1.0
        self.target_unsup_loss_weight = 1.0
        self.base_lr = 0.016
        self.momentum = 0.9
        self.weight_decay = 1e-4
        self.clip_norm = 10.0                      # pt/engine/trainer.py:385
        self.freeze_at = 2
        self.vgg_channels = [[64, 64], [128, 128], [256, 256, 256], [512, 512, 512], [512, 512, 512]]
        for k, v in kw.items():
            if not hasattr(self, k):
                raise KeyError(k)
            setattr(self, k, v)


# --------------------------------------------------------------------------------------------
# box arithmetic
# --------------------------------------------------------------------------------------------
def get_deltas(src, tgt, weights):
    """pt/modeling/box_regression.py:66-99 (note the +1e-9 inside the logs, :94-95)."""
    sw = src[:, 2] - src[:, 0]
    sh = src[:, 3] - src[:, 1]
    sx = src[:, 0] + 0.5 * sw
    sy = src[:, 1] + 0.5 * sh
    tw = tgt[:, 2] - tgt[:, 0]
    th = tgt[:, 3] - tgt[:, 1]
    tx = tgt[:, 0] + 0.5 * tw
    ty = tgt[:, 1] + 0.5 * th
    wx, wy, ww, wh = weights
    dx = wx * (tx - sx) / sw
    dy = wy * (ty - sy) / sh
    dw = ww * torch.log(tw / sw + 1e-9)
    dh = wh * torch.log(th / sh + 1e-9)
    assert bool((sw > 0).all()), "Input boxes to Box2BoxTransform are not valid!"
    return torch.stack((dx, dy, dw, dh), dim=1)


def apply_deltas(deltas, boxes, weights):
    """pt/modeling/box_regression.py:101-139 (deltas (N, k*4), stride-4 slicing, clamp dw/dh)."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = torch.clamp(deltas[:, 2::4] / ww, max=_SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3::4] / wh, max=_SCALE_CLAMP)
    pcx = dx * w[:, None] + cx[:, None]
    pcy = dy * h[:, None] + cy[:, None]
    pw = torch.exp(dw) * w[:, None]
    ph = torch.exp(dh) * h[:, None]
    out = torch.stack((pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph), dim=-1)
    return out.reshape(deltas.shape)


def gaussian_dist_pdf(val, mean, var, eps=1e-9):
    """pt/modeling/box_regression.py:33-35 (sigma constant 0.3 only in the normaliser)."""
    return torch.exp(-(val - mean) ** 2.0 / (var + eps) / 2.0) / torch.sqrt(2.0 * np.pi * (var + 0.3))


def pairwise_iou(b1, b2):
    """d2 v0.5 structures/boxes.py pairwise_iou: (M,4) x (R,4) -> (M,R)."""
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    wh = torch.min(b1[:, None, 2:], b2[:, 2:]) - torch.max(b1[:, None, :2], b2[:, :2])
    wh = wh.clamp(min=0)
    inter = wh.prod(dim=2)
    return torch.where(inter > 0, inter / (a1[:, None] + a2 - inter), torch.zeros(1, dtype=inter.dtype))


def clip_boxes(b, image_size):
    """d2 v0.5 Boxes.clip: x in [0,w], y in [0,h]."""
    h, w = image_size
    return torch.stack((b[:, 0].clamp(0, w), b[:, 1].clamp(0, h), b[:, 2].clamp(0, w), b[:, 3].clamp(0, h)), dim=-1)


def nonempty(b, threshold=0.0):
    return ((b[:, 2] - b[:, 0]) > threshold) & ((b[:, 3] - b[:, 1]) > threshold)


def matcher(iou, thresholds, labels, allow_low_quality):
    """d2 v0.5 modeling/matcher.py Matcher.__call__. Returns (matches int64 (R,), labels int8 (R,))."""
    if iou.numel() == 0:
        R = iou.shape[1]
        return iou.new_zeros(R, dtype=torch.int64), iou.new_full((R,), labels[0], dtype=torch.int8)
    vals, matches = iou.max(dim=0)
    out = matches.new_full(matches.size(), 1, dtype=torch.int8)
    th = [-float("inf")] + list(thresholds) + [float("inf")]
    for lab, lo, hi in zip(labels, th[:-1], th[1:]):
        out[(vals >= lo) & (vals < hi)] = lab
    if allow_low_quality:
        best_per_gt, _ = iou.max(dim=1)
        pred_with_best = (iou == best_per_gt[:, None]).any(dim=0)
        out[pred_with_best] = 1
    return matches, out


def _perm_from_prio(prio, n):
    """Sampling spec shared with the CUDA path: the random permutation of n items is the stable
    argsort of the first n entries of an injected priority vector (stands in for torch.randperm in
    d2 v0.5 subsample_labels; same distribution, injectable for parity)."""
    return torch.argsort(prio[:n], stable=True)


def subsample_labels(labels, num_samples, positive_fraction, bg_label, prio_pos, prio_neg):
    """d2 v0.5 modeling/sampling.py subsample_labels (two permutations, positives first)."""
    positive = torch.nonzero((labels != -1) & (labels != bg_label)).squeeze(1)
    negative = torch.nonzero(labels == bg_label).squeeze(1)
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    perm1 = _perm_from_prio(prio_pos, positive.numel())[:num_pos]
    perm2 = _perm_from_prio(prio_neg, negative.numel())[:num_neg]
    return positive[perm1], negative[perm2]


def nms(boxes, scores, thresh):
    """Greedy NMS (torchvision.ops.nms semantics: descending score, stable; suppress IoU > thresh).
    Returns kept indices in descending-score order."""
    if boxes.numel() == 0:
        return torch.zeros(0, dtype=torch.int64)
    b = boxes.detach().float().numpy()
    s = scores.detach().float().numpy()
    order = np.argsort(-s, kind="stable")
    b = b[order]
    areas = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    n = b.shape[0]
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(i)
        xx1 = np.maximum(b[i, 0], b[i + 1:, 0])
        yy1 = np.maximum(b[i, 1], b[i + 1:, 1])
        xx2 = np.minimum(b[i, 2], b[i + 1:, 2])
        yy2 = np.minimum(b[i, 3], b[i + 1:, 3])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        ovr = inter / (areas[i] + areas[i + 1:] - inter)
        suppressed[i + 1:] |= ovr > np.float32(thresh)
    return torch.from_numpy(order[np.asarray(keep, dtype=np.int64)])


def batched_nms(boxes, scores, idxs, thresh):
    """d2 v0.5 layers/nms.py batched_nms -> torchvision batched_nms, stated here as per-class NMS on
    un-shifted boxes (= torchvision `_batched_nms_vanilla`); result sorted by descending score."""
    if boxes.numel() == 0:
        return torch.zeros(0, dtype=torch.int64)
    keep_mask = torch.zeros_like(scores, dtype=torch.bool)
    for c in torch.unique(idxs):
        ci = torch.nonzero(idxs == c).squeeze(1)
        k = nms(boxes[ci], scores[ci], thresh)
        keep_mask[ci[k]] = True
    keep = torch.nonzero(keep_mask).squeeze(1)
    return keep[torch.argsort(-scores[keep], stable=True)]


# --------------------------------------------------------------------------------------------
# anchors
# --------------------------------------------------------------------------------------------
def default_cell_anchors(sizes, ratios):
    """d2 v0.5 DefaultAnchorGenerator.generate_cell_anchors (size-major, ratio-minor)."""
    out = []
    for size in sizes:
        area = size ** 2.0
        for r in ratios:
            w = math.sqrt(area / r)
            h = r * w
            out.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
    return torch.tensor(out, dtype=torch.float32)


def differentiable_cell_anchors(anchor_wh):
    """pt/modeling/anchor_generator.py:145-148."""
    return torch.stack([-anchor_wh[:, 0] / 2.0, -anchor_wh[:, 1] / 2.0, anchor_wh[:, 0] / 2.0,
                        anchor_wh[:, 1] / 2.0], -1)


def grid_anchors(cell, H, W, stride, offset):
    """pt/modeling/anchor_generator.py:108-122 + d2 v0.5 _create_grid_offsets."""
    sx = torch.arange(offset * stride, W * stride, step=stride, dtype=torch.float32)
    sy = torch.arange(offset * stride, H * stride, step=stride, dtype=torch.float32)
    yy, xx = torch.meshgrid(sy, sx, indexing="ij")
    xx = xx.reshape(-1)
    yy = yy.reshape(-1)
    shifts = torch.stack((xx, yy, xx, yy), dim=1)
    return (shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4)


class _GradZero(torch.autograd.Function):
    """pt/modeling/utils.py:47-58."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * 0.0


# --------------------------------------------------------------------------------------------
# proposal selection / pseudo-label filter
# --------------------------------------------------------------------------------------------
def find_top_rpn_proposals(proposals, logits, image_sizes, nms_thresh, pre_nms_topk, post_nms_topk,
                           min_box_size, training, sigma):
    """pt/modeling/proposal_generator/proposal_utils.py:27-154, single feature level.
    proposals (N,R,4), logits (N,R), sigma (N,R,4) raw sigma logits. Keeps the reference quirk at
    :94 -- sigma is taken from the FIRST k rows, not gathered by topk_idx.
    Tie order: the reference calls `logits.sort(descending=True)` (:87) without `stable`, so the order among EQUAL
    logits -- and through the :94 quirk the sigma row they are re-scored with -- is whatever the torch build does.
    Here (and in the CUDA path) ties are broken by ascending anchor index. fp32 ties do occur at full size
    (12 among the best 12 000 of 41 625 logits in tests/golden/pt_reference_config4_golden.pt, where they move 2 of
    the 2 000 proposals); none occur in the other fixtures, which this function reproduces exactly."""
    N, R = logits.shape
    k = min(R, pre_nms_topk)
    sl, idx = logits.sort(descending=True, dim=1, stable=True)
    topk_scores = sl[:, :k]
    topk_idx = idx[:, :k]
    topk_props = proposals[torch.arange(N)[:, None], topk_idx]
    topk_sigma = sigma[:, :k]
    results = []
    for n in range(N):
        boxes = topk_props[n]
        scores = topk_scores[n].clone()
        sg = topk_sigma[n]
        valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores)
        if not bool(valid.all()):
            if training:
                raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
            boxes, scores, sg = boxes[valid], scores[valid], sg[valid]
        boxes = clip_boxes(boxes, image_sizes[n])
        keep = nonempty(boxes, min_box_size)
        if int(keep.sum()) != boxes.shape[0]:
            boxes, scores = boxes[keep], scores[keep]
        sg = torch.sigmoid(sg[keep])
        scores = scores * (1 - sg.sum(-1) / 4.0)
        keep = batched_nms(boxes, scores, torch.zeros(boxes.shape[0], dtype=torch.int64), nms_thresh)
        keep = keep[:post_nms_topk]
        results.append(OInst(image_sizes[n], proposal_boxes=OBoxes(boxes[keep]), objectness_logits=scores[keep]))
    return results


def fast_rcnn_inference_single_image(boxes, scores, image_shape, score_thresh, nms_thresh, topk,
                                     cls_logits, deltas):
    """pt/modeling/roi_heads/fast_rcnn.py:34-120. boxes (R, K*8) = apply_deltas over all 8-vectors,
    scores (R,K+1) softmax, cls_logits (R,K+1) raw, deltas (R,K*8) raw."""
    R = boxes.shape[0]
    boxes = boxes.view(R, -1, 8)[..., :4].contiguous().view(R, -1)
    boxes_sigma = deltas.view(R, -1, 8)[..., -4:]
    valid = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not bool(valid.all()):
        boxes, scores, boxes_sigma, cls_logits = boxes[valid], scores[valid], boxes_sigma[valid], cls_logits[valid]
    K = boxes.shape[1] // 4
    scores = scores[:, :-1]
    boxes = clip_boxes(boxes.reshape(-1, 4), image_shape).view(-1, K, 4)
    boxes_sigma = boxes_sigma.reshape(-1, K, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if K == 1:
        boxes = boxes[filter_inds[:, 0], 0]
        boxes_sigma = boxes_sigma[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
        boxes_sigma = boxes_sigma[filter_mask]
    scores = scores[filter_mask]
    scores_logists = cls_logits[filter_inds[:, 0]]
    scores = scores * (1 - torch.sigmoid(boxes_sigma).sum(-1) / 4.0)
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk >= 0:
        keep = keep[:topk]
    res = OInst(image_shape, pred_boxes=OBoxes(boxes[keep]), scores=scores[keep],
                pred_classes=filter_inds[keep][:, 1], scores_logists=scores_logists[keep],
                boxes_sigma=boxes_sigma[keep])
    return res, filter_inds[keep][:, 0]


# --------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------
def rpn_losses(anchors, logits, gt_labels, deltas, gt_boxes, batch_per_image, weights):
    """pt/modeling/proposal_generator/rpn.py:191-255 + box_regression.py:142-176 (GUASSIAN branch).
    logits (N,R), deltas (N,R,8), gt_labels list of (R,) in {-1,0,1}, gt_boxes list of (R,4)."""
    N = len(gt_labels)
    lab = torch.stack(gt_labels)
    pos = lab == 1
    tgt = torch.stack([get_deltas(anchors, g, weights) for g in gt_boxes])
    var = torch.sigmoid(deltas[..., -4:])[pos]
    mean = deltas[..., :4][pos]
    g = gaussian_dist_pdf(mean, tgt[pos], var)
    loc = -torch.log(g + 1e-9).sum()
    valid = lab >= 0
    cls = F.binary_cross_entropy_with_logits(logits[valid], lab[valid].to(torch.float32), reduction="sum")
    norm = batch_per_image * N
    return {"loss_rpn_cls": cls / norm, "loss_rpn_loc": loc / norm}


def rpn_loss_unsupervised(logits, soft_labels, deltas, anchor_masks, matched_gt_boxes, matched_sigma,
                          anchors, efl, lam, tau, batch_per_image, weights):
    """pt/modeling/proposal_generator/rpn.py:257-361 (GUASSIAN). soft_labels list of (P_i,K+1) teacher
    logits, matched_sigma list of (P_i,4) teacher sigma logits, anchor_masks list of (R,) bool,
    matched_gt_boxes list of (R,4). Quirk kept: sigmoid([1-x, x]) at :299."""
    zt = torch.cat(soft_labels, 0)
    if efl:
        temp = torch.softmax(zt, -1)
        entropy = -(temp * torch.log(temp)).sum(-1)
        weight = (1 - entropy / math.log(temp.shape[-1])) ** lam[0]
    fg = zt.max(-1)[1] != (zt.shape[-1] - 1)
    gl = torch.softmax(zt / tau[0], -1).detach()
    gl = torch.stack([gl[:, -1], gl[:, :-1].sum(-1)], -1)
    masks = torch.stack(anchor_masks, 0)
    cls_out = logits[masks]
    cls_out = torch.sigmoid(torch.stack([1 - cls_out, cls_out], -1))
    cls_out = -torch.log(cls_out + 1e-9)
    if efl:
        gl = gl * weight.unsqueeze(-1)
    loss_cls = torch.sum(gl * cls_out)
    mean_p = torch.stack([get_deltas(anchors, k, weights) for k in matched_gt_boxes], 0)[masks]
    d = deltas[masks]
    sigma_p = torch.sigmoid(torch.cat(matched_sigma, 0)).detach()
    if efl:
        ent = 0.5 * torch.log(2 * np.pi * np.e * sigma_p)
        wb = (1 - ent / (0.5 * math.log(2 * np.pi * np.e))) ** lam[1]
    sigma_p = sigma_p * tau[1]
    sigma_q = torch.sigmoid(d[..., -4:])
    mean_q = d[..., :4]
    mean_p, sigma_p, mean_q, sigma_q = mean_p[fg], sigma_p[fg], mean_q[fg], sigma_q[fg]
    box = 0.5 * torch.log(sigma_q / sigma_p) - 0.5 + (sigma_p + (mean_q - mean_p) ** 2) / (2 * sigma_q)
    if efl:
        box = box * wb[fg]
    norm = batch_per_image * logits.shape[0]
    return {"loss_rpn_cls": loss_cls / norm, "loss_rpn_loc": box.sum() / norm}


def roi_cls_loss_unsupervised(zs, zt, efl, lam, tau):
    """pt/modeling/roi_heads/fast_rcnn.py:179-213."""
    zt = zt.detach()
    pq = -F.log_softmax(zs, -1)
    if efl:
        temp = F.softmax(zt, -1)
        entropy = -(temp * torch.log(temp)).sum(-1)
        weight = (1 - entropy / math.log(zt.shape[-1])) ** lam[0]
    sl = F.softmax(zt / tau[0], -1)
    if efl:
        sl = sl * weight.unsqueeze(-1)
    return torch.sum(sl * pq) / sl.shape[0]


def roi_box_loss_unsupervised(mean_q, sigma_q, mean_p, sigma_p, efl, lam, tau):
    """pt/modeling/roi_heads/fast_rcnn.py:215-263 (GUASSIAN)."""
    mean_p = mean_p.detach()
    sigma_p = torch.sigmoid(sigma_p).detach()
    if efl:
        ent = 0.5 * torch.log(2 * np.pi * np.e * sigma_p)
        weight = (1 - ent / (0.5 * math.log(2 * np.pi * np.e))) ** lam[1]
    sigma_p = sigma_p * tau[1]
    sigma_q = torch.sigmoid(sigma_q)
    loss = 0.5 * torch.log(sigma_q / sigma_p) - 0.5 + (sigma_p + (mean_q - mean_p) ** 2) / (2 * sigma_q)
    if efl:
        loss = loss * weight
    return loss.mean()


def roi_box_reg_loss(proposal_boxes, gt_boxes, pred_deltas, gt_classes, num_classes, weights):
    """pt/modeling/roi_heads/fast_rcnn.py:265-336 (GUASSIAN): class-selected 8-vector, Gaussian NLL,
    normalised by the number of sampled rois."""
    fg = torch.nonzero((gt_classes >= 0) & (gt_classes < num_classes)).squeeze(1)
    d = pred_deltas.view(-1, num_classes, 8)[fg, gt_classes[fg]]
    tgt = get_deltas(proposal_boxes[fg], gt_boxes[fg], weights)
    var = torch.sigmoid(d[..., -4:])
    g = gaussian_dist_pdf(d[..., :4], tgt, var)
    return -torch.log(g + 1e-9).sum() / max(gt_classes.numel(), 1.0)


# --------------------------------------------------------------------------------------------
# the detector (student or teacher)
# --------------------------------------------------------------------------------------------
class DefaultSampler:
    """Injected randomness: prio(tag, n) -> float32 priorities (see _perm_from_prio)."""

    def __init__(self, seed=0):
        self.gen = torch.Generator().manual_seed(seed)

    def prio(self, tag, n):
        return torch.rand(n, generator=self.gen)


class OracleRCNN(torch.nn.Module):
    """pt/modeling/meta_arch/rcnn.py:30-92 with its sub-modules inlined:
    backbone  pt/modeling/backbone/vgg.py:36-72,154-165,189-230
    rpn       pt/modeling/proposal_generator/rpn.py:44-154 (+ d2 v0.5 StandardRPNHead)
    roi heads pt/modeling/roi_heads/roi_heads.py:39-291 (+ d2 v0.5 ROIPooler, FastRCNNConvFCHead)
    Parameter names follow the reference state_dict."""

    def __init__(self, cfg: OracleCfg, seed=0):
        super().__init__()
        self.cfg = cfg
        g = torch.Generator().manual_seed(seed)
        P = torch.nn.Parameter
        cin = 3
        self.conv_names = []
        for bi, chans in enumerate(cfg.vgg_channels, start=1):
            for ci, cout in enumerate(chans, start=1):
                name = f"backbone.vgg_block{bi}.0.conv{ci}"
                fan_out = cout * 9
                w = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / fan_out)  # c2_msra_fill
                self._reg(name + ".weight", P(w))
                self._reg(name + ".bias", P(torch.zeros(cout)))
                self.conv_names.append((name, bi))
                cin = cout
        A = 9
        C = cin
        self._reg("proposal_generator.rpn_head.conv.weight", P(torch.randn(C, C, 3, 3, generator=g) * 0.01))
        self._reg("proposal_generator.rpn_head.conv.bias", P(torch.zeros(C)))
        self._reg("proposal_generator.rpn_head.objectness_logits.weight", P(torch.randn(A, C, 1, 1, generator=g) * 0.01))
        self._reg("proposal_generator.rpn_head.objectness_logits.bias", P(torch.zeros(A)))
        self._reg("proposal_generator.rpn_head.anchor_deltas.weight", P(torch.randn(A * 8, C, 1, 1, generator=g) * 0.01))
        self._reg("proposal_generator.rpn_head.anchor_deltas.bias", P(torch.zeros(A * 8)))
        if cfg.anchor_generator == "DifferentiableAnchorGenerator":
            self._reg("proposal_generator.anchor_generator.anchor_0", P(torch.tensor(cfg.anchor_wh, dtype=torch.float32)))
        K = cfg.num_classes
        fin = C * cfg.pooler_resolution ** 2
        for name, fi, fo in (("roi_heads.box_head.fc1", fin, cfg.fc_dim), ("roi_heads.box_head.fc2", cfg.fc_dim, cfg.fc_dim)):
            bound = math.sqrt(6.0 / fi)  # c2_xavier_fill = kaiming_uniform(a=1): bound sqrt(3/fan_in)*... (gain 1)
            bound = math.sqrt(3.0 / fi)
            self._reg(name + ".weight", P((torch.rand(fo, fi, generator=g) * 2 - 1) * bound))
            self._reg(name + ".bias", P(torch.zeros(fo)))
        self._reg("roi_heads.box_predictor.cls_score.weight", P(torch.randn(K + 1, cfg.fc_dim, generator=g) * 0.01))
        self._reg("roi_heads.box_predictor.cls_score.bias", P(torch.zeros(K + 1)))
        self._reg("roi_heads.box_predictor.bbox_pred.weight", P(torch.randn(K * 8, cfg.fc_dim, generator=g) * 0.001))
        self._reg("roi_heads.box_predictor.bbox_pred.bias", P(torch.zeros(K * 8)))
        # FREEZE_AT=2: blocks 1-2 frozen (vgg.py:175-180)
        for name, bi in self.conv_names:
            if bi <= cfg.freeze_at:
                self.p(name + ".weight").requires_grad_(False)
                self.p(name + ".bias").requires_grad_(False)
        self.register_buffer("pixel_mean", torch.tensor(cfg.pixel_mean).view(-1, 1, 1), persistent=False)
        self.register_buffer("pixel_std", torch.tensor(cfg.pixel_std).view(-1, 1, 1), persistent=False)
        self.sampler = DefaultSampler(0)

    # parameters live in a flat dict keyed by the reference's dotted names
    def _reg(self, name, p):
        self.register_parameter(name.replace(".", "__"), p)

    def p(self, name):
        return getattr(self, name.replace(".", "__"))

    def ref_state_dict(self):
        return {k.replace("__", "."): v for k, v in self.state_dict().items()}

    def load_ref_state_dict(self, sd):
        self.load_state_dict({k.replace(".", "__"): v for k, v in sd.items()})

    # ---------------------------------------------------------------- stages
    def preprocess_image(self, batched_inputs):
        """d2 v0.5 GeneralizedRCNN.preprocess_image + ImageList.from_tensors (pad to batch max)."""
        imgs = [(x["image"].float() - self.pixel_mean) / self.pixel_std for x in batched_inputs]
        sizes = [tuple(i.shape[-2:]) for i in imgs]
        H = max(s[0] for s in sizes)
        W = max(s[1] for s in sizes)
        out = imgs[0].new_zeros(len(imgs), 3, H, W)
        for i, im in enumerate(imgs):
            out[i, :, :im.shape[1], :im.shape[2]] = im
        return out, sizes

    def backbone(self, x, collect=None):
        """pt/modeling/backbone/vgg.py:65-72,154-165: conv+bias+ReLU, 2x2 maxpool after blocks 1-4."""
        last_block = 1
        for name, bi in self.conv_names:
            if bi != last_block:
                x = F.max_pool2d(x, 2, 2)
                last_block = bi
            x = F.relu(F.conv2d(x, self.p(name + ".weight"), self.p(name + ".bias"), padding=1))
            if collect is not None:
                collect[name] = x
        return x

    def rpn_head(self, feat):
        """d2 v0.5 StandardRPNHead.forward with box_dim 8 (rpn.py:44-55), then the layout permutes of
        rpn.py:97-113: logits (N, H*W*A), deltas (N, H*W*A, 8)."""
        t = F.relu(F.conv2d(feat, self.p("proposal_generator.rpn_head.conv.weight"),
                            self.p("proposal_generator.rpn_head.conv.bias"), padding=1))
        lg = F.conv2d(t, self.p("proposal_generator.rpn_head.objectness_logits.weight"),
                      self.p("proposal_generator.rpn_head.objectness_logits.bias"))
        dl = F.conv2d(t, self.p("proposal_generator.rpn_head.anchor_deltas.weight"),
                      self.p("proposal_generator.rpn_head.anchor_deltas.bias"))
        N = lg.shape[0]
        lg = lg.permute(0, 2, 3, 1).flatten(1)
        dl = dl.view(N, -1, 8, dl.shape[-2], dl.shape[-1]).permute(0, 3, 4, 1, 2).flatten(1, -2)
        return lg, dl

    def anchors(self, H, W, danchor):
        cfg = self.cfg
        if cfg.anchor_generator == "DifferentiableAnchorGenerator":
            cell = differentiable_cell_anchors(self.p("proposal_generator.anchor_generator.anchor_0"))
        else:
            cell = default_cell_anchors(cfg.anchor_sizes, cfg.anchor_ratios)
        a = grid_anchors(cell, H, W, cfg.stride, cfg.anchor_offset)
        if not danchor:
            a = _GradZero.apply(a)  # rpn.py:91-94
        return a

    def label_and_sample_anchors(self, anchors, gt_instances, use_ignore=False, use_soft_label=False):
        """pt/modeling/proposal_generator/rpn.py:363-448."""
        cfg = self.cfg
        anchors = anchors.detach()
        has_pseudo = gt_instances[0].has("pseudo_boxes")
        gt_labels, matched_gt, masks, msig = [], [], [], []
        for i, inst in enumerate(gt_instances):
            gt_boxes = _bt(inst.pseudo_boxes) if (use_ignore and has_pseudo) else _bt(inst.gt_boxes)
            iou = pairwise_iou(gt_boxes, anchors)
            midx, lab = matcher(iou, cfg.rpn_iou_thresholds, cfg.rpn_iou_labels, True)
            if has_pseudo and use_soft_label:
                amask = lab == 1
                sel = midx[amask]
                gt_labels.append(inst.scores_logists[sel])
                msig.append(inst.boxes_sigma[sel])
                masks.append(amask)
            else:
                R = lab.numel()
                pos, neg = subsample_labels(lab, cfg.rpn_batch_per_image, cfg.rpn_positive_fraction, 0,
                                            self.sampler.prio(("rpn_pos", i), R), self.sampler.prio(("rpn_neg", i), R))
                lab = lab.clone()
                lab.fill_(-1)
                lab[pos] = 1
                lab[neg] = 0
                gt_labels.append(lab)
            if gt_boxes.shape[0] == 0:
                matched_gt.append(torch.zeros_like(anchors))
            else:
                matched_gt.append(gt_boxes[midx])
        if has_pseudo and use_soft_label:
            return gt_labels, masks, matched_gt, msig
        return gt_labels, matched_gt

    def rpn(self, feat, image_sizes, gt_instances, compute_loss=True, branch="", danchor=False, training=True):
        """pt/modeling/proposal_generator/rpn.py:80-188."""
        cfg = self.cfg
        H, W = feat.shape[-2:]
        anchors = self.anchors(H, W, danchor)
        logits, deltas = self.rpn_head(feat)
        losses = {}
        if branch == "unsupervised":
            gl, masks, mgt, msig = self.label_and_sample_anchors(anchors, gt_instances, True, True)
            losses = rpn_loss_unsupervised(logits, gl, deltas, masks, mgt, msig, anchors, cfg.efl,
                                           cfg.efl_lambda, cfg.tau, cfg.rpn_batch_per_image, cfg.rpn_bbox_weights)
        elif training and compute_loss:
            gl, mgt = self.label_and_sample_anchors(anchors, gt_instances)
            losses = rpn_losses(anchors, logits, gl, deltas, mgt, cfg.rpn_batch_per_image, cfg.rpn_bbox_weights)
        with torch.no_grad():
            N, R = logits.shape
            props = apply_deltas(deltas[..., :4].reshape(-1, 4), anchors.detach().unsqueeze(0).expand(N, -1, -1).reshape(-1, 4),
                                 cfg.rpn_bbox_weights).view(N, -1, 4)
            ti = 0 if training else 1
            proposals = find_top_rpn_proposals(props, logits.detach(), image_sizes, cfg.rpn_nms_thresh,
                                               cfg.rpn_pre_nms_topk[ti], cfg.rpn_post_nms_topk[ti],
                                               cfg.rpn_min_size, training, deltas[..., 4:].detach())
        return proposals, losses, dict(logits=logits, deltas=deltas, anchors=anchors)

    def box_head(self, feat, boxes_per_image):
        """d2 v0.5 ROIPooler (ROIAlignV2: aligned=True, sampling_ratio 0, scale 1/16) +
        FastRCNNConvFCHead (flatten, fc1, ReLU, fc2, ReLU) + fast_rcnn.py:157-169 predictor."""
        from torchvision.ops import roi_align
        rois = torch.cat([torch.cat([torch.full((b.shape[0], 1), float(i)), b], 1) for i, b in enumerate(boxes_per_image)], 0)
        r = self.cfg.pooler_resolution
        x = roi_align(feat, rois, (r, r), 1.0 / self.cfg.stride, 0, True)
        x = x.flatten(1)
        x = F.relu(F.linear(x, self.p("roi_heads.box_head.fc1.weight"), self.p("roi_heads.box_head.fc1.bias")))
        x = F.relu(F.linear(x, self.p("roi_heads.box_head.fc2.weight"), self.p("roi_heads.box_head.fc2.bias")))
        scores = F.linear(x, self.p("roi_heads.box_predictor.cls_score.weight"), self.p("roi_heads.box_predictor.cls_score.bias"))
        deltas = F.linear(x, self.p("roi_heads.box_predictor.bbox_pred.weight"), self.p("roi_heads.box_predictor.bbox_pred.bias"))
        return scores, deltas

    def label_and_sample_proposals(self, proposals, targets, branch):
        """pt/modeling/roi_heads/roi_heads.py:192-291 + proposal_utils.py:157-224 + d2 v0.5 _sample_proposals."""
        cfg = self.cfg
        K = cfg.num_classes
        out = []
        for i, (prop, tgt) in enumerate(zip(proposals, targets)):
            pb = _bt(prop.proposal_boxes)
            if branch != "unsupervised":
                gtb = _bt(tgt.gt_boxes)
                pb = torch.cat([pb, gtb], 0)  # PROPOSAL_APPEND_GT
                iou = pairwise_iou(gtb, pb)
                midx, mlab = matcher(iou, [cfg.roi_iou_threshold], [0, 1], False)
                has_gt = gtb.shape[0] > 0
                if has_gt:
                    cls = tgt.gt_classes[midx].clone()
                    cls[mlab == 0] = K
                    cls[mlab == -1] = -1
                else:
                    cls = torch.zeros_like(midx) + K
                n = cls.numel()
                fg, bg = subsample_labels(cls, cfg.roi_batch_per_image, cfg.roi_positive_fraction, K,
                                          self.sampler.prio(("roi_pos", i), n), self.sampler.prio(("roi_neg", i), n))
                sidx = torch.cat([fg, bg], 0)
                gt_boxes = gtb[midx[sidx]] if has_gt else torch.zeros(sidx.numel(), 4)
                out.append(OInst(prop.image_size, proposal_boxes=OBoxes(pb[sidx]), gt_classes=cls[sidx],
                                 gt_boxes=OBoxes(gt_boxes), sampled_idx=sidx))
            else:
                psb = _bt(tgt.pseudo_boxes)
                iou = pairwise_iou(psb, pb)
                midx, mlab = matcher(iou, [cfg.roi_iou_threshold], [0, 1], False)
                sel = mlab == 1
                if psb.shape[0] == 0:
                    out.append(OInst(prop.image_size, proposal_boxes=OBoxes(pb[sel]), pseudo_boxes=OBoxes(psb),
                                     soft_label=tgt.scores_logists, boxes_sigma=tgt.boxes_sigma))
                else:
                    out.append(OInst(prop.image_size, proposal_boxes=OBoxes(pb[sel]), pseudo_boxes=OBoxes(psb[midx][sel]),
                                     soft_label=tgt.scores_logists[midx][sel], boxes_sigma=tgt.boxes_sigma[midx][sel]))
        return out

    def roi_heads(self, feat, proposals, targets, compute_loss=True, branch="", training=True):
        """pt/modeling/roi_heads/roi_heads.py:89-190."""
        cfg = self.cfg
        K = cfg.num_classes
        if training and compute_loss:
            proposals = self.label_and_sample_proposals(proposals, targets, branch)
        scores, deltas = self.box_head(feat, [_bt(p.proposal_boxes) for p in proposals])
        aux = dict(scores=scores, deltas=deltas, proposals=proposals)
        if branch == "unsupervised" and training:
            soft = torch.cat([p.soft_label for p in proposals])
            pseudo = torch.cat([_bt(p.pseudo_boxes) for p in proposals])
            losses = {"loss_cls": roi_cls_loss_unsupervised(scores, soft, cfg.efl, cfg.efl_lambda, cfg.tau)}
            sigma_p = torch.cat([p.boxes_sigma for p in proposals])
            pb = torch.cat([_bt(p.proposal_boxes) for p in proposals])
            mean_p = get_deltas(pb, pseudo, cfg.roi_bbox_weights)
            cls = soft.max(-1)[1]
            mask = cls != (soft.shape[-1] - 1)
            mq = deltas.view(-1, K, 8)[mask]
            sel = mq[torch.arange(mq.shape[0]), cls[mask]]
            losses["loss_box_reg"] = roi_box_loss_unsupervised(sel[:, :4], sel[:, -4:], mean_p[mask], sigma_p[mask],
                                                               cfg.efl, cfg.efl_lambda, cfg.tau)
            return proposals, losses, aux
        if training and compute_loss:
            gt_classes = torch.cat([p.gt_classes for p in proposals])
            pb = torch.cat([_bt(p.proposal_boxes) for p in proposals])
            gtb = torch.cat([_bt(p.gt_boxes) for p in proposals])
            losses = {"loss_cls": F.cross_entropy(scores, gt_classes, reduction="mean"),
                      "loss_box_reg": roi_box_reg_loss(pb, gtb, deltas, gt_classes, K, cfg.roi_bbox_weights)}
            return proposals, losses, aux
        # inference (teacher pseudo-label pass): fast_rcnn.py:338-409
        pb = torch.cat([_bt(p.proposal_boxes) for p in proposals])
        boxes = apply_deltas(deltas, pb, cfg.roi_bbox_weights)
        probs = F.softmax(scores, dim=-1)
        counts = [len(_bt(p.proposal_boxes)) for p in proposals]
        res = []
        for b, s, z, d, p in zip(boxes.split(counts), probs.split(counts), scores.split(counts), deltas.split(counts), proposals):
            r, _ = fast_rcnn_inference_single_image(b, s, p.image_size, cfg.roi_score_thresh_test,
                                                    cfg.roi_nms_thresh_test, cfg.detections_per_image, z, d)
            res.append(r)
        return res, {}, aux

    def forward(self, batched_inputs, branch="supervised", danchor=False, training=True, trace=None,
                proposals_override=None):
        """pt/modeling/meta_arch/rcnn.py:32-92. `proposals_override` (tests only) replaces the RPN
        proposals handed to the ROI heads, so that the ROI stage can be compared on identical inputs."""
        images, sizes = self.preprocess_image(batched_inputs)
        gt = [x["instances"] for x in batched_inputs] if "instances" in batched_inputs[0] else None
        feat = self.backbone(images, trace.setdefault("backbone", {}) if trace is not None else None)
        if branch == "supervised":
            props, pl, raux = self.rpn(feat, sizes, gt, training=training)
            if proposals_override is not None:
                props = proposals_override
            _, dl, haux = self.roi_heads(feat, props, gt, branch=branch, training=training)
            losses = dict(dl)
            losses.update(pl)
            out = (losses, [], [], None)
        elif branch == "unsup_data_weak":
            props, _, raux = self.rpn(feat, sizes, None, compute_loss=False, training=training)
            if proposals_override is not None:
                props = proposals_override
            roih, _, haux = self.roi_heads(feat, props, None, compute_loss=False, branch=branch, training=training)
            out = ({}, props, roih, (haux["scores"], haux["deltas"]))
        elif branch == "unsupervised":
            props, pl, raux = self.rpn(feat, sizes, gt, branch=branch, danchor=danchor, training=training)
            if proposals_override is not None:
                props = proposals_override
            _, dl, haux = self.roi_heads(feat, props, gt, branch=branch, training=training)
            losses = dict(dl)
            losses.update(pl)
            out = (losses, [], [], None)
        else:
            raise ValueError(branch)
        if trace is not None:
            trace.update(features=feat, rpn=raux, roi=haux, proposals=props)
        return out


# --------------------------------------------------------------------------------------------
# trainer step (pt/engine/trainer.py:263-392, 431-449, 557-603)
# --------------------------------------------------------------------------------------------
def resize_batch(data, ratios, pixel_mean):
    """pt/engine/trainer.py:557-590 with the random ratios injected. Returns new dict list."""
    out = []
    for d, ratio in zip(data, ratios):
        img = d["image"]
        h, w = img.shape[-2], img.shape[-1]
        d_h, d_w = int(h * ratio), int(w * ratio)
        x1 = int((w - d_w) / 2)
        y1 = int((h - d_h) / 2)
        bg = torch.zeros_like(img)
        bg += pixel_mean.view(-1, 1, 1).int().to(bg.dtype)
        bg[:, y1:y1 + d_h, x1:x1 + d_w] = F.interpolate(img.unsqueeze(0).float(), size=(d_h, d_w),
                                                        align_corners=False, mode="bilinear").squeeze(0).to(bg.dtype)
        nd = dict(d)
        nd["image"] = bg
        inst = d["instances"]
        ni = OInst(inst.image_size)
        for k in ("gt_boxes", "gt_classes", "pseudo_boxes", "scores_logists", "boxes_sigma"):
            if inst.has(k):
                v = getattr(inst, k)
                if k in ("gt_boxes", "pseudo_boxes"):
                    t = _bt(v).clone() * ratio
                    t[:, 0] += x1
                    t[:, 2] += x1
                    t[:, 1] += y1
                    t[:, 3] += y1
                    v = OBoxes(t)
                ni.set(k, v)
        nd["instances"] = ni
        out.append(nd)
    return out


def ema_update(teacher: OracleRCNN, student: OracleRCNN, keep_rate):
    """pt/engine/trainer.py:431-449."""
    with torch.no_grad():
        sd = student.state_dict()
        for k, v in teacher.state_dict().items():
            v.copy_(sd[k] * (1 - keep_rate) + v * keep_rate)


def clip_gradient(params, clip_norm):
    """pt/engine/trainer.py:592-603."""
    total = 0.0
    ps = [p for p in params if p.requires_grad and p.grad is not None]
    for p in ps:
        total = total + p.grad.norm() ** 2
    total = float(torch.sqrt(torch.as_tensor(total)))
    norm = clip_norm / max(total, clip_norm)
    for p in ps:
        p.grad.mul_(norm)
    return total


def run_step(student: OracleRCNN, teacher: OracleRCNN, optimizer, data, cfg: OracleCfg, ratios_unlabel,
             ratios_label, keep_rate=None, update_teacher=True):
    """One post-burn-in iteration, pt/engine/trainer.py:291-392. data = (label_q, label_k, unlabel_q,
    unlabel_k) lists of dicts. Returns the 8 losses as the reference logs them (floats, UNWEIGHTED: `metrics_dict =
    record_dict`, :379-381; SOURCE_LOSS_WEIGHT / TARGET_UNSUP_LOSS_WEIGHT only enter the sum that is back-propagated,
    :364-377). update_teacher=False: an iteration at which (iter - BURN_UP_STEP) % TEACHER_UPDATE_ITER != 0 (:296-301)."""
    label_q, label_k, unlabel_q, unlabel_k = data
    if update_teacher:
        ema_update(teacher, student, cfg.ema_keep_rate if keep_rate is None else keep_rate)
    with torch.no_grad():
        _, _, roih, _ = teacher(unlabel_k, branch="unsup_data_weak")
    pseudo = [OInst(r.image_size, pseudo_boxes=OBoxes(_bt(r.pred_boxes)), scores_logists=r.scores_logists,
                    boxes_sigma=r.boxes_sigma) for r in roih]  # trainer.py:179-228
    unlabel_q = [dict(d, instances=p) for d, p in zip(unlabel_q, pseudo)]
    unlabel_q = resize_batch(unlabel_q, ratios_unlabel, student.pixel_mean.flatten())
    label_q = resize_batch(label_q, ratios_label, student.pixel_mean.flatten())
    rec = {}
    l_sup, _, _, _ = student(label_q + label_k, branch="supervised")
    for k, v in l_sup.items():
        rec[k + "_sup"] = v
    l_un, _, _, _ = student(unlabel_q, branch="unsupervised", danchor=True)
    for k, v in l_un.items():
        rec[k + "_unsup"] = v
    total = sum(v * (cfg.source_loss_weight if k.endswith("_sup") else cfg.target_unsup_loss_weight)
                for k, v in rec.items())
    optimizer.zero_grad()
    total.backward()
    gn = clip_gradient(student.parameters(), cfg.clip_norm)
    optimizer.step()
    out = {k: float(v) for k, v in rec.items()}
    out["grad_norm"] = gn
    return out


def run_step_burn_in(student: OracleRCNN, optimizer, data, cfg: OracleCfg, ratios):
    """One source-only iteration (iter < BURN_UP_STEP), pt/engine/trainer.py:274-290,383-386: the strong and weak
    views of the labelled batch are concatenated (q first), ALL of them go through `resize` (one ratio each, in
    that order), one supervised forward, losses weighted 1.0, clip, SGD. The teacher is not touched. Returns the
    4 losses (floats, the reference's un-suffixed keys) + grad_norm."""
    label_q, label_k = data[0], data[1]
    batch = resize_batch(list(label_q) + list(label_k), ratios, student.pixel_mean.flatten())
    rec, _, _, _ = student(batch, branch="supervised")
    total = sum(v * 1.0 for v in rec.values())
    optimizer.zero_grad()
    total.backward()
    gn = clip_gradient(student.parameters(), cfg.clip_norm)
    optimizer.step()
    out = {k: float(v) for k, v in rec.items()}
    out["grad_norm"] = gn
    return out


def make_optimizer(model: OracleRCNN, cfg: OracleCfg):
    """d2 v0.5 build_optimizer: SGD(momentum, weight_decay) over all requires_grad params."""
    return torch.optim.SGD([p for p in model.parameters() if p.requires_grad], lr=cfg.base_lr,
                           momentum=cfg.momentum, weight_decay=cfg.weight_decay)


# --------------------------------------------------------------------------------------------
# synthetic data: seed 1234, uint8 uniform images, 12 GT boxes / image
# --------------------------------------------------------------------------------------------
def synthetic_batch(n, H, W, num_classes, seed, boxes_per_image=12, labelled=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        img = torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8)
        d = {"image": img, "height": H, "width": W}
        if labelled:
            wh = torch.rand(boxes_per_image, 2, generator=g) * (400 - 32) + 32
            wh[:, 0].clamp_(max=W - 2)
            wh[:, 1].clamp_(max=H - 2)
            cxy = torch.rand(boxes_per_image, 2, generator=g)
            x1 = cxy[:, 0] * (W - wh[:, 0])
            y1 = cxy[:, 1] * (H - wh[:, 1])
            boxes = torch.stack([x1, y1, x1 + wh[:, 0], y1 + wh[:, 1]], 1)
            cls = torch.randint(0, num_classes, (boxes_per_image,), generator=g)
            d["instances"] = OInst((H, W), gt_boxes=OBoxes(boxes), gt_classes=cls)
        out.append(d)
    return out
