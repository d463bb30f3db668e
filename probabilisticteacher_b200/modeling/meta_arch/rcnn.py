"""`GuassianGeneralizedRCNN` on the B200 path -- THE DROP-IN BOUNDARY.

Mirrors `pt/modeling/meta_arch/rcnn.py:30-92`: `model(batched_inputs, branch=..., danchor=...)` with
branch in {"supervised", "unsup_data_weak", "unsupervised"} returning
`(losses, proposals_rpn, proposals_roih, roi_predictions)`; the inherited detectron2 pieces
(GeneralizedRCNN.preprocess_image, device, pixel_mean) are restated here.

The four losses are returned as fp32 scalars attached to ONE autograd node (`_LossBundle`): calling
`.backward()` on any weighted sum of them (pt/engine/trainer.py:364-384) runs the hand-written
backward chain (head GEMM data/weight gradients -> ROIAlign backward -> conv data/weight gradients)
and accumulates into `param.grad`, which are views of the model's flat gradient arena.
"""
from collections import OrderedDict

import torch
from torch import nn

from ... import ops
from ..._lib import call, refresh_stream
from ...arena import ParamArena
from ...config import validate_cfg
from ...structures import Boxes, FreeInstances
from ..anchor_generator import ANCHOR_GENERATOR_REGISTRY
from ..backbone.vgg import build_vgg_backbone  # noqa: F401  (registers)
from ..proposal_generator.rpn import GuassianRPN  # noqa: F401
from ..registry import BACKBONE_REGISTRY, META_ARCH_REGISTRY, PROPOSAL_GENERATOR_REGISTRY, ROI_HEADS_REGISTRY
from ..roi_heads.roi_heads import GuassianROIHead  # noqa: F401


class _LossBundle(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hook, model, fctx, loss4):
        ctx.model = model
        ctx.fctx = fctx
        return tuple(loss4[i].clone() for i in range(4))

    @staticmethod
    def backward(ctx, g0, g1, g2, g3):
        g = [x.contiguous().to(torch.float32).reshape(1) for x in (g0, g1, g2, g3)]
        ctx.model._run_backward(ctx.fctx, g)
        return None, None, None, None


@META_ARCH_REGISTRY.register()
class GuassianGeneralizedRCNN(nn.Module):
    def __init__(self, cfg, device=None, loss_scale=1024.0, with_grads=True, precision="f16"):
        """precision: "f16x3" (split-fp16 operands = fp32-equivalent forward AND backward: the mode in which losses,
        logits and parameter gradients meet the 1e-3 parity with the reference's fp32 path) or "f16" (fp16 operands /
        fp32 accumulation: the mixed-precision throughput mode, 3x fewer tensor-core FLOPs)."""
        super().__init__()
        self.cfg = cfg
        dev = torch.device(device or cfg.MODEL.DEVICE)
        if dev.type != "cuda":
            raise RuntimeError("probabilisticteacher_b200 has no CPU path: a CUDA device is required")
        validate_cfg(cfg)  # refuse settings the reference honours but these kernels would silently ignore
        diff = cfg.MODEL.ANCHOR_GENERATOR.NAME == "DifferentiableAnchorGenerator"
        self.arena = ParamArena(num_classes=cfg.MODEL.ROI_HEADS.NUM_CLASSES, num_cell=9,
                                fc_dim=cfg.MODEL.ROI_BOX_HEAD.FC_DIM,
                                pooled=cfg.MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION,
                                freeze_at=cfg.MODEL.BACKBONE.FREEZE_AT, differentiable_anchors=diff, device=dev,
                                with_grads=with_grads)
        if precision not in ("f16", "f16x3"):
            raise ValueError(f"unknown precision {precision!r}")
        self.arena.precision = precision
        self.arena.pixel_mean = tuple(float(x) for x in cfg.MODEL.PIXEL_MEAN)
        self.arena.pixel_std = tuple(float(x) for x in cfg.MODEL.PIXEL_STD)
        self.loss_scale = float(loss_scale)
        self.backbone = BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, self.arena, self.loss_scale)
        anchor_gen = ANCHOR_GENERATOR_REGISTRY.get(cfg.MODEL.ANCHOR_GENERATOR.NAME)(cfg, self.arena)
        self.proposal_generator = PROPOSAL_GENERATOR_REGISTRY.get(cfg.MODEL.PROPOSAL_GENERATOR.NAME)(
            cfg, self.arena, anchor_gen, self.loss_scale)
        self.roi_heads = ROI_HEADS_REGISTRY.get(cfg.MODEL.ROI_HEADS.NAME)(cfg, self.arena, self.loss_scale)
        self.register_buffer("pixel_mean", torch.tensor(cfg.MODEL.PIXEL_MEAN, device=dev).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(cfg.MODEL.PIXEL_STD, device=dev).view(-1, 1, 1), False)
        self._mean = [float(x) for x in cfg.MODEL.PIXEL_MEAN]
        self._std = [float(x) for x in cfg.MODEL.PIXEL_STD]
        # parameters: views into the arena under the reference's names
        self._param_names = []
        for name, v, g, trainable in self.arena.exposed_parameters():
            p = nn.Parameter(v, requires_grad=trainable)
            self._param_names.append(name)
            self.register_parameter(name.replace(".", "__"), p)
        self._hook = torch.zeros(1, device=dev, requires_grad=True)
        self._pending = 0
        self.prio_generator = None
        self.prio_override = None  # tests inject {tag: (prio_pos, prio_neg)}
        self._hw_cache = {}
        self.backward_stream = None
        self.heads_backward_hook = None
        self.launch_count = 0

    # ------------------------------------------------------------------ nn.Module surface
    @property
    def device(self):
        return self.pixel_mean.device

    def named_parameters(self, prefix="", recurse=True, remove_duplicate=True):
        for n, p in super().named_parameters(prefix, recurse, remove_duplicate):
            yield n.replace("__", "."), p

    def state_dict(self, *args, **kwargs):
        """Reference layout / reference key names (see arena.py)."""
        return self.arena.state_dict()

    def load_state_dict(self, sd, strict=True):
        """strict: a missing key raises KeyError; otherwise it is skipped. Returns torch's
        (missing_keys, unexpected_keys) named tuple (what fvcore's Checkpointer._load_model reads)."""
        from torch.nn.modules.module import _IncompatibleKeys
        missing, unexpected = self.arena.load_state_dict(sd, strict=strict)
        return _IncompatibleKeys(missing, unexpected)

    def init_synthetic(self, seed=0):
        return self.arena.init_synthetic(seed)

    def attach_grads(self):
        """Makes every trainable parameter's .grad a view of the flat gradient arena."""
        pairs = getattr(self, "_grad_pairs", None)
        if pairs is None:
            pairs = []
            for (name, v, g, trainable), (_, p) in zip(self.arena.exposed_parameters(), super().named_parameters()):
                if trainable and g is not None:
                    pairs.append((p, g))
            self._grad_pairs = pairs
        for p, g in pairs:
            if p.grad is not g:
                p.grad = g

    def zero_grad(self, set_to_none=False):
        if self.arena.grads is not None:
            self.arena.grads.zero_()
        self.attach_grads()

    # ------------------------------------------------------------------ inputs
    def _prio(self, tag, N, L):
        if self.prio_override is not None and tag in self.prio_override:
            return self.prio_override[tag]
        dev = self.device
        g = self.prio_generator
        return (torch.rand(N, L, device=dev, generator=g), torch.rand(N, L, device=dev, generator=g))

    def preprocess_image(self, batched_inputs):
        """d2 GeneralizedRCNN.preprocess_image + ImageList.from_tensors: normalise, zero-pad to the batch
        max; fused with the first VGG conv (bias + ReLU). Returns (FlatAct C=64, image_sizes, img_hw)."""
        dev = self.device
        imgs = [x["image"] for x in batched_inputs]
        sizes = [tuple(i.shape[-2:]) for i in imgs]
        H = max(s[0] for s in sizes)
        W = max(s[1] for s in sizes)
        same = all(s == (H, W) for s in sizes)
        imgs = [i if i.is_cuda else i.to(dev, non_blocking=True) for i in imgs]
        if same:
            batch = torch.stack(imgs) if len(imgs) > 1 else imgs[0].unsqueeze(0)
        else:
            batch = torch.zeros(len(imgs), 3 * H * W, dtype=torch.uint8, device=dev)
            for k, i in enumerate(imgs):
                batch[k, :i.numel()] = i.reshape(-1)
        batch = batch.contiguous()
        key = tuple(sizes)
        cached = self._hw_cache.get(key)
        if cached is None:  # (cached so that a CUDA-graph capture never sees a host->device copy here)
            hw_i = torch.tensor(sizes, dtype=torch.int32).to(dev)
            cached = (hw_i, hw_i.to(torch.float32))
            self._hw_cache[key] = cached
        hw_i, img_hw = cached
        name0 = self.arena.conv_specs[0][0]
        if self.arena.precision == "f16x3":
            if ops.X3_CONV1_TC[0]:
                if self.arena.conv1_x3 is None:
                    self.arena.pack_x3(dgrad=self.arena.dgrad_half is not None)
                pack, table, alpha = self.arena.conv1_x3
                act = ops.conv1_u8_x3_tc(batch.view(len(imgs), -1), hw_i, H, W, pack, table, alpha)
            else:  # fp32 CUDA-core version (cross-check)
                act = ops.conv1_u8_x3(batch.view(len(imgs), -1), hw_i, H, W, self._mean, self._std,
                                      self.arena.view(name0 + ".weight").view(64, 27), self.arena.view(name0 + ".bias"))
            return act, sizes, img_hw
        act = ops.conv1_u8(batch.view(len(imgs), -1), hw_i, H, W, self._mean, self._std, self.arena.conv1_half,
                           self.arena.view(name0 + ".bias"))
        return act, sizes, img_hw

    def _targets(self, instances):
        """list[FreeInstances] -> padded device tensors (no host sync for device-resident pseudo labels)."""
        dev = self.device
        N = len(instances)
        t = {}
        if instances[0].has("gt_boxes") and instances[0].gt_boxes.tensor.is_cuda:
            # device-resident ground truth (fixed capacity + device count): no host traffic
            cap = max(len(i.gt_boxes) for i in instances)
            t["gt_boxes"] = torch.stack([i.gt_boxes.tensor for i in instances]) if all(
                len(i.gt_boxes) == cap for i in instances) else None
            assert t["gt_boxes"] is not None, "device-resident gt must share one capacity"
            t["gt_classes"] = torch.stack([i.gt_classes.to(torch.int32) for i in instances])
            cnts = []
            for i in instances:
                c = i.valid_count() if isinstance(i, FreeInstances) else None
                cnts.append(c.reshape(1).to(torch.int32) if c is not None else
                            torch.full((1,), len(i.gt_boxes), dtype=torch.int32, device=dev))
            t["gt_count"] = torch.cat(cnts)
        elif instances[0].has("gt_boxes"):
            cnt = [len(i.gt_boxes) for i in instances]
            cap = max(16, (max(cnt) + 15) // 16 * 16)
            gb = torch.zeros(N, cap, 4, dtype=torch.float32)
            gc = torch.zeros(N, cap, dtype=torch.int32)
            for k, i in enumerate(instances):
                if cnt[k]:
                    gb[k, :cnt[k]] = i.gt_boxes.tensor.detach().to("cpu", torch.float32)
                    gc[k, :cnt[k]] = i.gt_classes.detach().to("cpu", torch.int32)
            t["gt_boxes"] = gb.to(dev, non_blocking=True)
            t["gt_classes"] = gc.to(dev, non_blocking=True)
            t["gt_count"] = torch.tensor(cnt, dtype=torch.int32).to(dev, non_blocking=True)
        if instances[0].has("pseudo_boxes"):
            cap = max(len(i.pseudo_boxes) for i in instances)
            cap = max(16, cap)

            def pad(x, shape):
                x = x.to(dev, torch.float32)
                if x.shape[0] == shape[0]:
                    return x
                o = torch.zeros(shape, dtype=torch.float32, device=dev)
                o[:x.shape[0]] = x
                return o
            K1 = instances[0].scores_logists.shape[-1]
            t["pseudo_boxes"] = torch.stack([pad(i.pseudo_boxes.tensor, (cap, 4)) for i in instances])
            t["scores_logists"] = torch.stack([pad(i.scores_logists, (cap, K1)) for i in instances])
            t["boxes_sigma"] = torch.stack([pad(i.boxes_sigma, (cap, 4)) for i in instances])
            cnts = []
            for i in instances:
                c = i.valid_count() if isinstance(i, FreeInstances) else None
                cnts.append(c.reshape(1).to(torch.int32) if c is not None else
                            torch.tensor([len(i.pseudo_boxes)], dtype=torch.int32, device=dev))
            t["pseudo_count"] = torch.cat(cnts)
        return t

    # ------------------------------------------------------------------ forward
    def forward(self, batched_inputs, branch="supervised", danchor=False, norm=False, targets_provider=None):
        """`targets_provider` (optional, trainer-internal): callable returning the list of instances once
        the backbone has been issued -- lets the pseudo labels of the unsupervised branch arrive from
        another stream while this branch's backbone already runs."""
        refresh_stream()
        if not self.training:
            return self.inference(batched_inputs)
        if norm and "instances" in batched_inputs[0]:
            # rcnn.py:37-38 calls `self.preprocess_image_norm`, which the reference defines nowhere: same failure here
            raise AttributeError("'GuassianGeneralizedRCNN' object has no attribute 'preprocess_image_norm'")
        act, sizes, img_hw = self.preprocess_image(batched_inputs)
        need_grad = branch in ("supervised", "unsupervised") and torch.is_grad_enabled()
        feats, records = self.backbone(act, save=need_grad)
        feat = feats["vgg_block5"]
        targets = None
        if targets_provider is not None:
            targets = self._targets(targets_provider())
            refresh_stream()
        elif "instances" in batched_inputs[0]:
            targets = self._targets([x["instances"] for x in batched_inputs])
        if branch == "supervised":
            props, l_rpn, rctx = self.proposal_generator(feat, img_hw, targets, prio=self._prio)
            _, l_roi, hctx = self.roi_heads(feat, props, img_hw, targets, branch=branch, prio=self._prio)
        elif branch == "unsup_data_weak":
            props, _, rctx = self.proposal_generator(feat, img_hw, None, compute_loss=False)
            res, _, hctx = self.roi_heads(feat, props, img_hw, None, compute_loss=False, branch=branch)
            return {}, self._proposal_instances(props, sizes), self._roih_instances(res, sizes), \
                (hctx["scores"], hctx["deltas"])
        elif branch == "unsupervised":
            props, l_rpn, rctx = self.proposal_generator(feat, img_hw, targets, branch=branch, danchor=danchor)
            _, l_roi, hctx = self.roi_heads(feat, props, img_hw, targets, branch=branch)
        else:
            raise ValueError(f"unknown branch {branch!r}")
        loss4 = torch.cat([l_roi, l_rpn])
        if need_grad:
            fctx = dict(records=records, rpn=rctx, roi=hctx, feat=feat)
            self._pending += 1
            l0, l1, l2, l3 = _LossBundle.apply(self._hook, self, fctx, loss4)
        else:
            l0, l1, l2, l3 = loss4.unbind(0)
        losses = OrderedDict(loss_cls=l0, loss_box_reg=l1, loss_rpn_cls=l2, loss_rpn_loc=l3)
        self._last_ctx = dict(rpn=rctx, roi=hctx, feat=feat, props=props)
        return losses, [], [], None

    def inference(self, batched_inputs, do_postprocess=True):
        """Eval-mode path (rcnn.py:33-34 -> d2 `GeneralizedRCNN.inference`): same kernels with the test-time top-k
        (6000 / 1000), then `detector_postprocess` to the "height" / "width" of each input dict (exact-length
        instances, one host sync per image). do_postprocess=False returns the fixed-capacity device instances."""
        with torch.no_grad():
            act, sizes, img_hw = self.preprocess_image(batched_inputs)
            feats, _ = self.backbone(act, save=False)
            feat = feats["vgg_block5"]
            props, _, _ = self.proposal_generator(feat, img_hw, None, compute_loss=False, training=False)
            res, _, _ = self.roi_heads(feat, props, img_hw, None, compute_loss=False, training=False)
        instances = self._roih_instances(res, sizes)
        if do_postprocess:
            from ..postprocessing import postprocess_batch
            return postprocess_batch(instances, batched_inputs, sizes)
        return instances

    def _proposal_instances(self, props, sizes):
        out = []
        for n, s in enumerate(sizes):
            inst = FreeInstances(s)
            inst.proposal_boxes = Boxes(props["boxes"][n])
            inst.objectness_logits = props["scores"][n]
            inst._count = props["count"][n]
            out.append(inst)
        return out

    def _roih_instances(self, res, sizes):
        out = []
        for n, s in enumerate(sizes):
            inst = FreeInstances(s)
            inst.pred_boxes = Boxes(res["pred_boxes"][n])
            inst.scores = res["scores"][n]
            inst.pred_classes = res["pred_classes"][n]
            inst.scores_logists = res["scores_logists"][n]
            inst.boxes_sigma = res["boxes_sigma"][n]
            inst._count = res["count"][n]
            out.append(inst)
        return out

    # ------------------------------------------------------------------ backward
    def _run_backward(self, fctx, g):
        """g = [g_loss_cls, g_loss_box_reg, g_loss_rpn_cls, g_loss_rpn_loc] device scalars."""
        bs = self.backward_stream
        if bs is not None:
            # autograd replays a node on the stream its forward ran on; the trainer's concurrent step
            # wants the whole backward on ONE stream (after the join of the forward branches)
            cur = torch.cuda.current_stream()
            bs.wait_stream(cur)
            with torch.cuda.stream(bs):
                self._run_backward_impl(fctx, g)
            cur.wait_stream(bs)
            return
        self._run_backward_impl(fctx, g)

    def _run_backward_impl(self, fctx, g):
        refresh_stream()
        self.attach_grads()
        feat = fctx["feat"]
        dfeat_roi = self.roi_heads.backward(fctx["roi"], g[0], g[1])
        dfeat_rpn = self.proposal_generator.backward(fctx["rpn"], g[2], g[3])
        if self.heads_backward_hook is not None:
            # every gradient outside the VGG backbone (box head, predictor, RPN head, anchors) is final for this
            # pass: the trainer records an event here and all-reduces that arena suffix while the backbone
            # backward still runs
            self.heads_backward_hook()
        dz = torch.empty_like(feat.t)
        if self.arena.precision == "f16x3":  # both paths arrive un-masked in fp32; feat / dz are triples
            C = self.arena.C
            call("ptb200_add_mask_f16x3", dfeat_rpn, dfeat_roi, feat.t, dz, dz.numel() // (3 * C), C)
        else:
            call("ptb200_add_mask_f16", dfeat_rpn.t, dfeat_roi, 1.0, feat.t, dz, dz.numel())
        self.backbone.backward(fctx["records"], ops.FlatAct(dz, feat.H, feat.W))
        self._pending -= 1
        fctx.clear()


def build_model(cfg, device=None, **kw):
    """`DefaultTrainer.build_model` as used at pt/engine/trainer.py:79,85."""
    return META_ARCH_REGISTRY.get(cfg.MODEL.META_ARCHITECTURE)(cfg, device=device, **kw)
