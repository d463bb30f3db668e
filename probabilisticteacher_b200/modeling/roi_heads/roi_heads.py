"""Gaussian ROI head on the B200 path: the counterpart of `pt/modeling/roi_heads/roi_heads.py`
(GuassianROIHead: _init_box_head :53-87, forward :89-117, _forward_box :119-190,
label_and_sample_proposals :192-255, _sample_proposals_unsup :257-291) with detectron2's ROIPooler
(ROIAlignV2) and FastRCNNConvFCHead (2 x FC 1024) inlined as kernels."""
import torch
from torch import nn

from ... import ops
from ..._lib import call
from .. import sampling
from ..registry import ROI_HEADS_REGISTRY
from .fast_rcnn import GuassianFastRCNNOutputLayers


@ROI_HEADS_REGISTRY.register()
class GuassianROIHead(nn.Module):
    def __init__(self, cfg, arena, loss_scale):
        super().__init__()
        self.cfg = cfg
        self.arena = arena
        self.loss_scale = loss_scale
        r = cfg.MODEL.ROI_HEADS
        self.num_classes = r.NUM_CLASSES
        self.batch_size_per_image = r.BATCH_SIZE_PER_IMAGE
        self.positive_fraction = r.POSITIVE_FRACTION
        self.iou_threshold = r.IOU_THRESHOLDS[0]
        self.proposal_append_gt = r.PROPOSAL_APPEND_GT
        self.pooler_resolution = cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION
        self.pooler_scale = 1.0 / 16
        self.box_predictor = GuassianFastRCNNOutputLayers(cfg, arena)

    # ------------------------------------------------------------------ box head
    def _box_head(self, feat, rois, counts, cap):
        ar = self.arena
        if ar.precision == "f16x3":
            return self._box_head_x3(feat, rois, counts, cap)
        x0 = ops.roi_align_fwd(feat, rois, counts, cap, self.pooler_scale, self.pooler_resolution)
        rows = x0.shape[0]
        p = "roi_heads.box_head."
        w1 = ar.hview(p + "fc1.weight").view(ar.fc_dim, -1)
        tiles = ((rows + 127) // 128) * (ar.fc_dim // 256)
        ksplit = max(1, min(8, 148 // max(tiles, 1)))
        seg = (counts, cap)  # 128-row tiles without a live roi are skipped by the GEMMs
        h1 = ops.gemm_tn(x0.view(1, rows, -1), w1, epi=ops.EPI_BIAS_RELU, bias=ar.view(p + "fc1.bias"), ksplit=ksplit,
                         seg=seg)
        h2 = ops.gemm_tn(h1, ar.hview(p + "fc2.weight"), epi=ops.EPI_BIAS_RELU, bias=ar.view(p + "fc2.bias"), seg=seg)
        scores, deltas = self.box_predictor(h2.view(rows, -1), seg=seg)
        return x0, h1.view(rows, -1), h2.view(rows, -1), scores, deltas

    def _box_head_x3(self, feat, rois, counts, cap):
        """Split-fp16 fp32-equivalent precision of the box head."""
        ar = self.arena
        x0 = ops.roi_align_fwd_x3(feat, rois, counts, cap, self.pooler_scale, self.pooler_resolution)
        rows = x0.shape[0]
        seg = (counts, cap)
        p = "roi_heads.box_head."
        w3, alpha = ar.x3view(p + "fc1.weight")
        h1 = ops.gemm_tn_x3(x0.view(1, rows, -1), w3, alpha, bias=ar.view(p + "fc1.bias"), seg=seg)
        w3, alpha = ar.x3view(p + "fc2.weight")
        h2 = ops.gemm_tn_x3(h1, w3, alpha, bias=ar.view(p + "fc2.bias"), seg=seg)
        scores, deltas = self.box_predictor(h2.view(rows, -1), seg=seg)
        return x0, h1.view(rows, -1), h2.view(rows, -1), scores, deltas

    def forward(self, feat: ops.FlatAct, proposals, img_hw, targets=None, compute_loss=True, branch="",
                training=True, prio=None):
        """proposals: dict(boxes [N,P,4], scores, count). Returns (result, loss2 or None, ctx)."""
        N = feat.t.shape[0]
        K = self.num_classes
        if training and compute_loss:
            if branch == "unsupervised":
                m = sampling.roi_match_unsup(targets["pseudo_boxes"], targets["scores_logists"],
                                             targets["boxes_sigma"], targets["pseudo_count"], proposals["boxes"],
                                             proposals["count"], self.iou_threshold)
                cap = proposals["boxes"].shape[1]
                x0, h1, h2, scores, deltas = self._box_head(feat, m["rois"], m["count"], cap)
                loss2, ds, dd = self.box_predictor.losses_unsupervised(scores, deltas, m, N, cap)
                sel = m
            else:
                assert self.proposal_append_gt
                L = proposals["boxes"].shape[1] + targets["gt_boxes"].shape[1]
                pp, pn = prio("roi", N, L)
                sel = sampling.roi_label_and_sample(targets["gt_boxes"], targets["gt_classes"], targets["gt_count"],
                                                    proposals["boxes"], proposals["count"], K, self.iou_threshold,
                                                    self.batch_size_per_image, self.positive_fraction, pp, pn)
                cap = self.batch_size_per_image
                x0, h1, h2, scores, deltas = self._box_head(feat, sel["rois"], sel["count"], cap)
                loss2, ds, dd = self.box_predictor.losses(scores, deltas, sel, N, cap)
            ctx = dict(feat=feat, rois=sel["rois"], counts=sel["count"], cap=cap, x0=x0, h1=h1, h2=h2, dscores=ds,
                       ddeltas=dd, N=N, scores=scores, deltas=deltas, sel=sel)
            return sel, loss2, ctx
        cap = proposals["boxes"].shape[1]
        x0, h1, h2, scores, deltas = self._box_head(feat, proposals["boxes"], proposals["count"], cap)
        res = self.box_predictor.inference(scores, deltas, proposals["boxes"], proposals["count"], img_hw, N, cap)
        return res, None, dict(scores=scores, deltas=deltas)

    # ------------------------------------------------------------------ backward
    def backward(self, ctx, g_cls, g_box):
        """Accumulates box-head gradients; returns d(loss)/d(feat) as fp32 [N, H*(W+1), C] (un-masked,
        scaled by the loss scale)."""
        ar = self.arena
        K = self.num_classes
        S = self.loss_scale
        inv = 1.0 / S
        rows = ctx["x0"].shape[0]
        dev = ctx["x0"].device
        fc = ar.fc_dim
        seg = (ctx["counts"], ctx["cap"])
        if ar.precision == "f16x3":
            return self._backward_x3(ctx, g_cls, g_box)
        dpred = torch.empty(rows, 128, dtype=torch.float16, device=dev)
        call("ptb200_pack_grad2_f16", ctx["dscores"], K + 1, ctx["ddeltas"], 8 * K, g_cls, g_box, S, rows, 128, dpred)
        p = "roi_heads.box_predictor."
        ops.wgrad(dpred.view(1, rows, 128), ctx["h2"].view(1, rows, fc), ar.gview(p + "_heads.weight"), scale=inv,
                  bias_out=ar.gview(p + "_heads.bias"), seg=seg)
        dz2 = ops.gemm_tn(dpred.view(1, rows, 128), ar.dgrad_half["pred"], epi=ops.EPI_MASK, aux=ctx["h2"], seg=seg)
        p = "roi_heads.box_head."
        ops.wgrad(dz2, ctx["h1"].view(1, rows, fc), ar.gview(p + "fc2.weight"), scale=inv,
                  bias_out=ar.gview(p + "fc2.bias"), seg=seg)
        dz1 = ops.gemm_tn(dz2, ar.dgrad_half["fc2"], epi=ops.EPI_MASK, aux=ctx["h1"], seg=seg)
        fin = ctx["x0"].shape[1]
        ops.wgrad(dz1, ctx["x0"].view(1, rows, fin), ar.gview(p + "fc1.weight").view(fc, fin), scale=inv,
                  bias_out=ar.gview(p + "fc1.bias"), seg=seg)
        dx0 = ops.gemm_tn(dz1, ar.dgrad_half["fc1"], epi=ops.EPI_BIAS, seg=seg)
        return ops.roi_align_bwd(dx0.view(rows, fin), ctx["feat"], ctx["rois"], ctx["counts"], ctx["cap"],
                                 self.pooler_scale, self.pooler_resolution)

    def _backward_x3(self, ctx, g_cls, g_box):
        """f16x3 precision of `backward` (the reference's fp32 autograd): saved activations and output gradients are
        triples, the fc1 data gradient is stored in fp32 for the ROIAlign backward."""
        ar = self.arena
        K = self.num_classes
        S = self.loss_scale
        inv = 1.0 / S
        rows = ctx["x0"].shape[0]
        fc = ar.fc_dim
        seg = (ctx["counts"], ctx["cap"])
        dpred = ops.pack_grad2_x3(ctx["dscores"], K + 1, ctx["ddeltas"], 8 * K, g_cls, g_box, S, rows, 128)
        dpred = dpred.view(1, rows, 384)
        h2, h1, x0 = ctx["h2"].view(1, rows, 3 * fc), ctx["h1"].view(1, rows, 3 * fc), ctx["x0"]
        fin = x0.shape[1] // 3
        p = "roi_heads.box_predictor."
        ops.wgrad_x3(dpred, h2, ar.gview(p + "_heads.weight"), m_total=128, n_total=fc, scale=inv,
                     bias_out=ar.gview(p + "_heads.bias"), seg=seg)
        wd3, alpha = ar.dgrad_x3["pred"]
        dz2 = ops.gemm_tn_x3(dpred, wd3, alpha, epi=ops.EPI_SPLIT3_MASK, aux=h2, seg=seg)
        p = "roi_heads.box_head."
        ops.wgrad_x3(dz2, h1, ar.gview(p + "fc2.weight"), m_total=fc, n_total=fc, scale=inv,
                     bias_out=ar.gview(p + "fc2.bias"), seg=seg)
        wd3, alpha = ar.dgrad_x3["fc2"]
        dz1 = ops.gemm_tn_x3(dz2, wd3, alpha, epi=ops.EPI_SPLIT3_MASK, aux=h1, seg=seg)
        ops.wgrad_x3(dz1, x0.view(1, rows, 3 * fin), ar.gview(p + "fc1.weight").view(fc, fin), m_total=fc, n_total=fin,
                     scale=inv, bias_out=ar.gview(p + "fc1.bias"), seg=seg)
        wd3, alpha = ar.dgrad_x3["fc1"]
        dx0 = ops.gemm_tn_x3(dz1, wd3, alpha, epi=ops.EPI_F32_STORE, seg=seg)
        return ops.roi_align_bwd_f32(dx0.view(rows, fin), ctx["feat"], ctx["rois"], ctx["counts"], ctx["cap"],
                                     self.pooler_scale, self.pooler_resolution)
