"""Gaussian RPN on the B200 path: the counterpart of `pt/modeling/proposal_generator/rpn.py`
(GuassianRPNHead :44-55, GuassianRPN.forward :80-154, predict_proposals :156-188, losses :191-255,
loss_rpn_unsupervised :257-361, label_and_sample_anchors :363-448).

Differences in form, not in semantics: head outputs stay in the flat NHWC row layout (the reference's
permutes at :97-113 are free here), labelling / sampling / losses are fused device kernels with
fixed-capacity buffers, and backward is explicit (`backward()` returns the gradient w.r.t. the
backbone feature map)."""
import torch
from torch import nn

from ... import ops
from ..._lib import call
from .. import sampling
from ..registry import PROPOSAL_GENERATOR_REGISTRY, RPN_HEAD_REGISTRY
from .proposal_utils import find_top_rpn_proposals


@RPN_HEAD_REGISTRY.register()
class GuassianRPNHead(nn.Module):
    """3x3 conv + ReLU, then one fused 1x1 GEMM producing A objectness logits and A*8 (mu, sigma)
    regression outputs per location (box_dim doubled as in rpn.py:50-55)."""

    def __init__(self, arena):
        super().__init__()
        self.arena = arena

    def forward(self, feat: ops.FlatAct):
        ar = self.arena
        C, A = ar.C, ar.A
        p = "proposal_generator.rpn_head."
        n_valid = A * 9
        n_total = (n_valid + 15) // 16 * 16
        if ar.precision == "f16x3":
            w3, alpha = ar.x3view(p + "conv.weight")
            t = ops.conv3x3_x3(feat, w3, alpha, ar.view(p + "conv.bias"))
            w3, alpha = ar.x3view(p + "_heads.weight")
            logits, deltas = ops.gemm_tn_x3(t.t, w3, alpha, epi=ops.EPI_F32_SPLIT, bias=ar.view(p + "_heads.bias"),
                                            split=A, n_valid=n_valid, n_total=n_total, bn=n_total)
            return t, logits, deltas
        t = ops.conv3x3(feat, ar.hview(p + "conv.weight").view(C, 9 * C), ar.view(p + "conv.bias"), relu=True)
        logits, deltas = ops.gemm_tn(t.t, ar.hview(p + "_heads.weight"), epi=ops.EPI_F32_SPLIT,
                                     bias=ar.view(p + "_heads.bias"), split=A, n_valid=n_valid, n_total=n_total,
                                     bn=n_total)
        return t, logits, deltas


@PROPOSAL_GENERATOR_REGISTRY.register()
class GuassianRPN(nn.Module):
    def __init__(self, cfg, arena, anchor_generator, loss_scale):
        super().__init__()
        self.cfg = cfg
        self.arena = arena
        self.anchor_generator = anchor_generator
        self.rpn_head = RPN_HEAD_REGISTRY.get(cfg.MODEL.RPN.HEAD_NAME)(arena)
        r = cfg.MODEL.RPN
        self.iou_thresholds = tuple(r.IOU_THRESHOLDS)
        self.batch_size_per_image = r.BATCH_SIZE_PER_IMAGE
        self.positive_fraction = r.POSITIVE_FRACTION
        self.pre_nms_topk = {True: r.PRE_NMS_TOPK_TRAIN, False: r.PRE_NMS_TOPK_TEST}
        self.post_nms_topk = {True: r.POST_NMS_TOPK_TRAIN, False: r.POST_NMS_TOPK_TEST}
        self.nms_thresh = r.NMS_THRESH
        self.min_box_size = float(cfg.MODEL.PROPOSAL_GENERATOR.MIN_SIZE)
        self.loss_weight = float(r.LOSS_WEIGHT)
        assert tuple(r.BBOX_REG_WEIGHTS) == (1.0, 1.0, 1.0, 1.0), "kernels assume MODEL.RPN.BBOX_REG_WEIGHTS = 1"
        self.loss_scale = loss_scale
        self.nonfinite_flag = torch.zeros(1, dtype=torch.int32, device=arena.device)

    # ------------------------------------------------------------------ forward
    def forward(self, feat: ops.FlatAct, img_hw, targets=None, compute_loss=True, branch="", danchor=False,
                training=True, prio=None):
        """targets: dict with padded device tensors (see GuassianGeneralizedRCNN._targets).
        Returns (proposals dict(boxes, scores, count), loss2 tensor or None, ctx)."""
        cfg = self.cfg
        N = feat.t.shape[0]
        H, W, A = feat.H, feat.W, self.arena.A
        R = H * W * A
        anchors = self.anchor_generator(H, W)
        t, logits, deltas = self.rpn_head(feat)
        ctx = dict(feat=feat, t=t, N=N, H=H, W=W, branch=branch, danchor=danchor)
        loss2 = None
        dev = feat.t.device
        rows = N * H * (W + 1)
        norm = self.loss_weight / (self.batch_size_per_image * N)
        # MODEL.RPN.LOSS_WEIGHT multiplies the supervised losses only (rpn.py:136-141); loss_rpn_unsupervised is unweighted
        norm_unsup = 1.0 / (self.batch_size_per_image * N)
        if branch == "unsupervised":
            matched, labels = sampling.rpn_match(targets["pseudo_boxes"], targets["pseudo_count"], anchors, N,
                                                 self.iou_thresholds[0], self.iou_thresholds[1])
            loss2 = torch.empty(2, dtype=torch.float32, device=dev)
            dl = torch.empty(rows, A, dtype=torch.float32, device=dev)
            dd = torch.empty(rows, A * 8, dtype=torch.float32, device=dev)
            da = torch.empty(A, 2, dtype=torch.float32, device=dev) if (danchor and self.anchor_generator.differentiable) else None
            u = cfg.UNSUPNET
            call("ptb200_rpn_loss_unsup", logits, A, deltas, A * 8, labels, matched, targets["pseudo_boxes"],
                 targets["scores_logists"], targets["boxes_sigma"], targets["pseudo_boxes"].shape[1], anchors, N, H, W,
                 A, targets["scores_logists"].shape[2], int(bool(u.EFL)), float(u.EFL_LAMBDA[0]),
                 float(u.EFL_LAMBDA[1]), float(u.TAU[0]), float(u.TAU[1]), norm_unsup, loss2, dl, dd, da)
            ctx.update(dlogits=dl, ddeltas=dd, danchor_wh=da)
        elif training and compute_loss:
            matched, labels = sampling.rpn_match(targets["gt_boxes"], targets["gt_count"], anchors, N,
                                                 self.iou_thresholds[0], self.iou_thresholds[1])
            pp, pn = prio("rpn", N, R)
            sampled = sampling.rpn_subsample(labels, self.batch_size_per_image, self.positive_fraction, pp, pn)
            loss2 = torch.empty(2, dtype=torch.float32, device=dev)
            dl = torch.empty(rows, A, dtype=torch.float32, device=dev)
            dd = torch.empty(rows, A * 8, dtype=torch.float32, device=dev)
            call("ptb200_rpn_loss_sup", logits, A, deltas, A * 8, sampled, matched, targets["gt_boxes"],
                 targets["gt_boxes"].shape[1], anchors, N, H, W, A, norm, loss2, dl, dd)
            ctx.update(dlogits=dl, ddeltas=dd, danchor_wh=None, labels=sampled)
        # predict_proposals (rpn.py:156-188): no gradient flows through the proposals
        boxes, scores, count = find_top_rpn_proposals(
            logits, deltas, anchors, N, H, W, A, img_hw, self.nms_thresh, self.pre_nms_topk[training],
            self.post_nms_topk[training], self.min_box_size, self.nonfinite_flag)
        ctx.update(logits=logits, deltas=deltas, anchors=anchors)
        return dict(boxes=boxes, scores=scores, count=count), loss2, ctx

    def raise_if_nonfinite(self):
        """The reference checks every image's decoded proposals on the host and raises in training
        (`proposal_utils.py:117-122`). Here the decode kernel drops non-finite candidates and sets a device flag
        instead (no per-image synchronisation); this reads the flag (ONE host sync), clears it and raises the
        reference's error. Called by `PTrainer.check_finite` every few iterations, not inside the step."""
        if int(self.nonfinite_flag.item()) != 0:
            self.nonfinite_flag.zero_()
            raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")

    # ------------------------------------------------------------------ backward
    def backward(self, ctx, g_cls, g_loc):
        """g_cls / g_loc: device scalars (upstream gradients of loss_rpn_cls / loss_rpn_loc).
        Accumulates parameter gradients; returns d(loss)/d(feat) as a ReLU-masked fp16 FlatAct
        (scaled by the loss scale)."""
        ar = self.arena
        C, A = ar.C, ar.A
        S = self.loss_scale
        inv = 1.0 / S
        feat, t = ctx["feat"], ctx["t"]
        N = ctx["N"]
        rows = N * feat.H * (feat.W + 1)
        dev = feat.t.device
        if ar.precision == "f16x3":
            return self._backward_x3(ctx, g_cls, g_loc)
        dhead = torch.empty(rows, 128, dtype=torch.float16, device=dev)
        call("ptb200_pack_grad2_f16", ctx["dlogits"], A, ctx["ddeltas"], A * 8, g_cls, g_loc, S, rows, 128, dhead)
        p = "proposal_generator.rpn_head."
        ops.wgrad(dhead.view(1, rows, 128), t.t.view(1, rows, C), ar.gview(p + "_heads.weight"), scale=inv,
                  bias_out=ar.gview(p + "_heads.bias"))
        dzt = ops.gemm_tn(dhead.view(1, rows, 128), ar.dgrad_half["rpn_heads"], epi=ops.EPI_MASK, aux=t.t)
        dzt = ops.FlatAct(dzt.view(N, -1, C), feat.H, feat.W)
        ops.conv3x3_wgrad(dzt, feat, ar.gview(p + "conv.weight").view(C, 9 * C), scale=inv,
                          bias_out=ar.gview(p + "conv.bias"))
        dfeat = ops.conv3x3(dzt, ar.dgrad_half["rpn_conv"], None, aux=feat.t)
        if ctx.get("danchor_wh") is not None:
            call("ptb200_axpy_dev", g_loc, 1.0, ctx["danchor_wh"],
                 ar.gview("proposal_generator.anchor_generator.anchor_0"), A * 2)
        return dfeat

    def _backward_x3(self, ctx, g_cls, g_loc):
        """f16x3 precision: same chain over triples; returns d(loss)/d(feat) as an UN-masked fp32 tensor
        [N, H*(W+1), C] (scaled by the loss scale; the ReLU mask of the backbone output is applied when the RPN and
        ROI paths are summed)."""
        ar = self.arena
        C, A = ar.C, ar.A
        S = self.loss_scale
        inv = 1.0 / S
        feat, t = ctx["feat"], ctx["t"]
        N = ctx["N"]
        rows = N * feat.H * (feat.W + 1)
        dhead = ops.pack_grad2_x3(ctx["dlogits"], A, ctx["ddeltas"], A * 8, g_cls, g_loc, S, rows, 128)
        p = "proposal_generator.rpn_head."
        ops.wgrad_x3(dhead.view(1, rows, 384), t.t.view(1, rows, 3 * C), ar.gview(p + "_heads.weight"), m_total=128,
                     n_total=C, scale=inv, bias_out=ar.gview(p + "_heads.bias"))
        wd3, alpha = ar.dgrad_x3["rpn_heads"]
        dzt = ops.gemm_tn_x3(dhead.view(1, rows, 384), wd3, alpha, epi=ops.EPI_SPLIT3_MASK, aux=t.t.view(1, rows, 3 * C))
        dzt = ops.FlatAct(dzt.view(N, -1, 3 * C), feat.H, feat.W)
        ops.conv3x3_wgrad_x3(dzt, feat, ar.gview(p + "conv.weight").view(C, 9 * C), C, C, scale=inv,
                             bias_out=ar.gview(p + "conv.bias"))
        wd3, alpha = ar.dgrad_x3["rpn_conv"]
        dfeat = ops.conv3x3_dgrad_x3(dzt, wd3, alpha)
        if ctx.get("danchor_wh") is not None:
            call("ptb200_axpy_dev", g_loc, 1.0, ctx["danchor_wh"],
                 ar.gview("proposal_generator.anchor_generator.anchor_0"), A * 2)
        return dfeat
