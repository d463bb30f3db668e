"""Gaussian Fast R-CNN output layers on the B200 path: the counterpart of
`pt/modeling/roi_heads/fast_rcnn.py` -- fast_rcnn_inference(_single_image) :34-141 (the teacher's
pseudo-label filter), GuassianFastRCNNOutputLayers :145-409 (8-dim bbox_pred, supervised Gaussian NLL,
unsupervised soft-label CE / KL losses, inference)."""
import torch
from torch import nn

from ... import ops
from ..._lib import call


def fast_rcnn_inference(scores, deltas, props, prop_count, img_hw, N, cap, num_classes, score_thresh, nms_thresh,
                        topk_per_image, bbox_weights):
    """scores fp32 [N*cap, K+1], deltas fp32 [N*cap, 8K], props fp32 [N, cap, 4]. Returns dict of
    fixed-capacity tensors [N, topk, ...] + count int32 [N] (fast_rcnn.py:34-120 per image)."""
    dev = scores.device
    K = num_classes
    cb = torch.empty(N, cap * K, 4, dtype=torch.float32, device=dev)
    cs = torch.empty(N, cap * K, dtype=torch.float32, device=dev)
    keys = torch.empty(N, cap * K, dtype=torch.int32, device=dev)
    vals = torch.empty(N, cap * K, dtype=torch.int32, device=dev)
    cc = torch.empty(N, dtype=torch.int32, device=dev)
    call("ptb200_roi_infer_candidates", scores, deltas, props, prop_count, N, cap, K, img_hw, float(score_thresh),
         list(bbox_weights), cb, cs, keys, vals, cc)
    ops.segmented_sort(keys, vals)
    keep_idx, keep_count = ops.nms(cb, vals, cc, nms_thresh, topk_per_image, class_mod=K)
    T = topk_per_image
    out = dict(pred_boxes=torch.empty(N, T, 4, dtype=torch.float32, device=dev),
               scores=torch.empty(N, T, dtype=torch.float32, device=dev),
               pred_classes=torch.empty(N, T, dtype=torch.int64, device=dev),
               scores_logists=torch.empty(N, T, K + 1, dtype=torch.float32, device=dev),
               boxes_sigma=torch.empty(N, T, 4, dtype=torch.float32, device=dev),
               src_roi=torch.empty(N, T, dtype=torch.int32, device=dev), count=keep_count)
    call("ptb200_roi_infer_gather", cb, cs, scores, deltas, vals, keep_idx, keep_count, N, cap, K, T,
         out["pred_boxes"], out["scores"], out["pred_classes"], out["scores_logists"], out["boxes_sigma"],
         out["src_roi"])
    return out


class GuassianFastRCNNOutputLayers(nn.Module):
    """cls_score (K+1) and bbox_pred (K*8 = mu,sigma per class) as ONE GEMM over a padded [128, 1024]
    weight block with an fp32 split epilogue."""

    def __init__(self, cfg, arena):
        super().__init__()
        self.cfg = cfg
        self.arena = arena
        self.num_classes = cfg.MODEL.ROI_HEADS.NUM_CLASSES
        self.box2box_weights = tuple(cfg.MODEL.ROI_BOX_HEAD.BBOX_REG_WEIGHTS)
        self.test_score_thresh = cfg.MODEL.ROI_HEADS.SCORE_THRESH_TEST
        self.test_nms_thresh = cfg.MODEL.ROI_HEADS.NMS_THRESH_TEST
        self.test_topk_per_image = cfg.TEST.DETECTIONS_PER_IMAGE
        self.model_type = cfg.UNSUPNET.MODEL_TYPE
        if self.model_type != "GUASSIAN":
            raise ValueError("only UNSUPNET.MODEL_TYPE == 'GUASSIAN' is on the hot path")

    def forward(self, h2, seg=None):
        """h2: fp16 [rows, fc_dim] -> (scores fp32 [rows, K+1], deltas fp32 [rows, 8K])."""
        ar = self.arena
        K = self.num_classes
        n_valid = (K + 1) + 8 * K
        n_total = (n_valid + 15) // 16 * 16
        rows = h2.shape[0]
        p = "roi_heads.box_predictor."
        if ar.precision == "f16x3":
            w3, alpha = ar.x3view(p + "_heads.weight")
            s, d = ops.gemm_tn_x3(h2.view(1, rows, -1), w3, alpha, epi=ops.EPI_F32_SPLIT,
                                  bias=ar.view(p + "_heads.bias"), split=K + 1, n_valid=n_valid, n_total=n_total,
                                  bn=n_total, seg=seg)
            return s.view(rows, K + 1), d.view(rows, 8 * K)
        s, d = ops.gemm_tn(h2.view(1, rows, -1), ar.hview(p + "_heads.weight"), epi=ops.EPI_F32_SPLIT,
                           bias=ar.view(p + "_heads.bias"), split=K + 1, n_valid=n_valid, n_total=n_total, bn=n_total,
                           seg=seg)
        return s.view(rows, K + 1), d.view(rows, 8 * K)

    def losses(self, scores, deltas, sampled, N, cap):
        """Supervised: mean CE + Gaussian NLL (fast_rcnn.py:265-336). Returns (loss2, dscores, ddeltas)."""
        dev = scores.device
        K = self.num_classes
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        ds = torch.empty(N * cap, K + 1, dtype=torch.float32, device=dev)
        dd = torch.empty(N * cap, 8 * K, dtype=torch.float32, device=dev)
        call("ptb200_roi_loss_sup", scores, deltas, sampled["gt_classes"], sampled["rois"], sampled["gt_boxes"],
             sampled["count"], N, cap, K, list(self.box2box_weights), loss2, ds, dd)
        return loss2, ds, dd

    def losses_unsupervised(self, scores, deltas, matched, N, cap):
        """cls_loss_unsupervised :179-213 + box_reg_loss_unsupervised :215-263 with the class-selected
        8-vector of roi_heads.py:146-164."""
        dev = scores.device
        K = self.num_classes
        u = self.cfg.UNSUPNET
        loss2 = torch.empty(2, dtype=torch.float32, device=dev)
        totals = torch.empty(2, dtype=torch.int32, device=dev)
        ds = torch.empty(N * cap, K + 1, dtype=torch.float32, device=dev)
        dd = torch.empty(N * cap, 8 * K, dtype=torch.float32, device=dev)
        call("ptb200_roi_loss_unsup", scores, deltas, matched["soft_label"], matched["boxes_sigma"], matched["rois"],
             matched["pseudo_boxes"], matched["count"], N, cap, K, int(bool(u.EFL)), float(u.EFL_LAMBDA[0]),
             float(u.EFL_LAMBDA[1]), float(u.TAU[0]), float(u.TAU[1]), list(self.box2box_weights), totals, loss2, ds,
             dd)
        return loss2, ds, dd

    def inference(self, scores, deltas, props, prop_count, img_hw, N, cap):
        return fast_rcnn_inference(scores, deltas, props, prop_count, img_hw, N, cap, self.num_classes,
                                   self.test_score_thresh, self.test_nms_thresh, self.test_topk_per_image,
                                   self.box2box_weights)
