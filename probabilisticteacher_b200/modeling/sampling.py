"""Matching and sub-sampling on device (detectron2 v0.5 Matcher / subsample_labels as used at
`pt/modeling/proposal_generator/rpn.py:414-433` and `pt/modeling/roi_heads/roi_heads.py:207-225`).

Randomness is injected: the permutation of n candidates is the stable argsort of the first n entries
of a priority vector (uniform [0,1) floats drawn with torch on the device, or supplied by a test)."""
import torch

from .._lib import call
from .. import ops


def rpn_match(gt_boxes, gt_count, anchors, N, iou_lo, iou_hi):
    """gt_boxes fp32 [N, cap, 4], gt_count int32 [N], anchors fp32 [R, 4] -> (matched int32 [N,R],
    labels int32 [N,R] in {-1,0,1} before sub-sampling)."""
    dev = anchors.device
    R = anchors.shape[0]
    cap = gt_boxes.shape[1]
    max_iou = torch.empty(N, R, dtype=torch.float32, device=dev)
    best = torch.empty(N, cap, dtype=torch.int32, device=dev)
    matched = torch.empty(N, R, dtype=torch.int32, device=dev)
    labels = torch.empty(N, R, dtype=torch.int32, device=dev)
    call("ptb200_rpn_match", gt_boxes, gt_count, cap, anchors, R, N, float(iou_lo), float(iou_hi), max_iou, best,
         matched, labels)
    return matched, labels


def _select(labels, seg_len, bg_label, prio_pos, prio_neg):
    """labels int32 [N, L]. Returns (pos_list, neg_list, perm [2N, L], counts [N, 2])."""
    N, L = labels.shape
    dev = labels.device
    pos_list = torch.empty(N, L, dtype=torch.int32, device=dev)
    neg_list = torch.empty(N, L, dtype=torch.int32, device=dev)
    counts = torch.empty(N, 2, dtype=torch.int32, device=dev)
    call("ptb200_compact_pos_neg", labels, L, seg_len, L, N, bg_label, pos_list, neg_list, counts)
    keys = torch.empty(2 * N, L, dtype=torch.int32, device=dev)
    perm = torch.empty(2 * N, L, dtype=torch.int32, device=dev)
    call("ptb200_prio_keys", prio_pos, prio_neg, L, counts, N, keys, perm)
    ops.segmented_sort(keys, perm, seg_len=counts.view(-1), max_len=L)
    return pos_list, neg_list, perm, counts


def rpn_subsample(labels, batch_per_image, positive_fraction, prio_pos, prio_neg):
    """d2 RPN._subsample_labels: int32 [N,R] labels -> int8 [N,R] with -1 except the sampled 1 / 0."""
    N, R = labels.shape
    pos_list, neg_list, perm, counts = _select(labels, None, 0, prio_pos, prio_neg)
    out = torch.empty(N, R, dtype=torch.int8, device=labels.device)
    call("ptb200_rpn_sample_apply", pos_list, neg_list, perm, R, counts, N, R, batch_per_image,
         int(batch_per_image * positive_fraction), out)
    return out


def roi_label_and_sample(gt_boxes, gt_classes, gt_count, props, prop_count, num_classes, iou_thr,
                         batch_per_image, positive_fraction, prio_pos, prio_neg):
    """roi_heads.py:192-255 (supervised branch). gt_* are [N, gcap, ...], props fp32 [N, pcap, 4].
    Returns dict(rois [N,B,4], gt_classes int32 [N,B], gt_boxes [N,B,4], count int32 [N], src int32 [N,B])."""
    N, pcap = props.shape[:2]
    gcap = gt_boxes.shape[1]
    L = pcap + gcap
    dev = props.device
    cls = torch.empty(N, L, dtype=torch.int32, device=dev)
    matched = torch.empty(N, L, dtype=torch.int32, device=dev)
    cand = torch.empty(N, dtype=torch.int32, device=dev)
    call("ptb200_roi_label", gt_boxes, gt_classes, gt_count, gcap, props, prop_count, pcap, N, num_classes,
         float(iou_thr), cls, matched, cand)
    pos_list, neg_list, perm, counts = _select(cls, cand, num_classes, prio_pos, prio_neg)
    B = batch_per_image
    rois = torch.empty(N, B, 4, dtype=torch.float32, device=dev)
    ocls = torch.empty(N, B, dtype=torch.int32, device=dev)
    ogt = torch.empty(N, B, 4, dtype=torch.float32, device=dev)
    ocount = torch.empty(N, dtype=torch.int32, device=dev)
    osrc = torch.empty(N, B, dtype=torch.int32, device=dev)
    call("ptb200_roi_sample_apply", pos_list, neg_list, perm, L, counts, cls, matched, gt_boxes, gt_count, gcap,
         props, prop_count, pcap, N, B, int(B * positive_fraction), num_classes, rois, ocls, ogt, ocount, osrc)
    return dict(rois=rois, gt_classes=ocls, gt_boxes=ogt, count=ocount, src=osrc)


def roi_match_unsup(pseudo_boxes, pseudo_logits, pseudo_sigma, pseudo_count, props, prop_count, iou_thr):
    """roi_heads.py:257-291: keep (in order) the proposals whose best pseudo box has IoU >= thr."""
    N, pcap = props.shape[:2]
    scap = pseudo_boxes.shape[1]
    K1 = pseudo_logits.shape[2]
    dev = props.device
    rois = torch.zeros(N, pcap, 4, dtype=torch.float32, device=dev)
    rois[..., 2:] = 1.0
    ops_ = torch.zeros(N, pcap, 4, dtype=torch.float32, device=dev)
    logits = torch.zeros(N, pcap, K1, dtype=torch.float32, device=dev)
    sigma = torch.zeros(N, pcap, 4, dtype=torch.float32, device=dev)
    count = torch.empty(N, dtype=torch.int32, device=dev)
    call("ptb200_roi_match_unsup", pseudo_boxes, pseudo_logits, pseudo_sigma, pseudo_count, scap, props, prop_count,
         pcap, N, K1, float(iou_thr), rois, ops_, logits, sigma, count)
    return dict(rois=rois, pseudo_boxes=ops_, soft_label=logits, boxes_sigma=sigma, count=count)
