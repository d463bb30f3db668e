"""Device-side proposal selection: the B200 counterpart of
`pt/modeling/proposal_generator/proposal_utils.py:27-154` (find_top_rpn_proposals) and
`:157-224` (add_ground_truth_to_proposals, fused into ptb200_roi_label)."""
import torch

from ..._lib import call
from ... import ops


def find_top_rpn_proposals(logits, deltas, anchors, N, H, W, num_cell, img_hw, nms_thresh, pre_nms_topk,
                           post_nms_topk, min_box_size, nonfinite_flag):
    """logits: fp32 [N, H*(W+1), A]; deltas: fp32 [N, H*(W+1), A*8] (head outputs, flat rows);
    anchors: fp32 [R, 4]; img_hw: fp32 [N, 2] device. Returns (boxes [N, post, 4], scores [N, post],
    count int32 [N]); rows >= count are zero. No host synchronisation."""
    dev = logits.device
    R = H * W * num_cell
    k = min(R, pre_nms_topk)
    keys = torch.empty(N, R, dtype=torch.int32, device=dev)
    vals = torch.empty(N, R, dtype=torch.int32, device=dev)
    call("ptb200_rpn_make_keys", logits, logits.shape[2], N, H, W, num_cell, keys, vals)
    ops.segmented_sort(keys, vals)
    boxes = torch.empty(N, k, 4, dtype=torch.float32, device=dev)
    scores = torch.empty(N, k, dtype=torch.float32, device=dev)
    keys2 = torch.empty(N, k, dtype=torch.int32, device=dev)
    vals2 = torch.empty(N, k, dtype=torch.int32, device=dev)
    valid = torch.empty(N, dtype=torch.int32, device=dev)
    call("ptb200_rpn_topk_decode", vals, R, logits, logits.shape[2], deltas, deltas.shape[2], anchors, N, H, W,
         num_cell, k, img_hw, float(min_box_size), boxes, scores, keys2, vals2, valid, nonfinite_flag)
    ops.segmented_sort(keys2, vals2)
    keep_idx, keep_count = ops.nms(boxes, vals2, valid, nms_thresh, post_nms_topk)
    out_boxes = torch.empty(N, post_nms_topk, 4, dtype=torch.float32, device=dev)
    out_scores = torch.empty(N, post_nms_topk, dtype=torch.float32, device=dev)
    call("ptb200_rpn_gather", boxes, scores, k, vals2, k, keep_idx, keep_count, post_nms_topk, N, out_boxes,
         out_scores)
    return out_boxes, out_scores, keep_count
