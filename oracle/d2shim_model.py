"""Functional detectron2 v0.5 base classes for running the reference's OWN model classes end to end
-- TEST INFRASTRUCTURE, NOT PRODUCT CODE (only oracle/make_golden_model.py uses it, in this container).

`oracle/d2shim.py` registers empty base classes, enough to import the reference's files and to call their
methods one by one. This module replaces those placeholders with working restatements of the v0.5 base classes
(GeneralizedRCNN, Backbone, RPN, StandardRPNHead, DefaultAnchorGenerator, StandardROIHeads, ROIPooler,
FastRCNNConvFCHead, FastRCNNOutputLayers), so that the reference's `GuassianGeneralizedRCNN`, `VGG`,
`GuassianRPN`, `GuassianRPNHead`, `DifferentiableAnchorGenerator`, `GuassianROIHead` and
`GuassianFastRCNNOutputLayers` can be CONSTRUCTED and their `forward` methods executed unmodified: the
orchestration inside the reference's classes (rcnn.py:30-92, rpn.py:80-188, roi_heads.py:89-291,
fast_rcnn.py:150-409) is then pinned by the reference's own code, not by a restatement.

Each base class follows the detectron2 v0.5 source behaviour listed in SURVEY.md 8c (detectron2 itself is not
installable here). Random sub-sampling is injected: `PRIO.prio((tag, image_index), n)` supplies the priority
vectors whose stable argsort replaces `torch.randperm` (same convention as oracle/pt_oracle.py).
"""
import sys

import torch
import torch.nn.functional as F
from torch import nn

from oracle import d2shim


class _Prio:
    """Injected sampling priorities: set `provider` (callable(tag, n) -> tensor) and call reset() per forward."""

    def __init__(self):
        self.provider = None
        self.counters = {}

    def reset(self):
        self.counters = {}

    def next_index(self, group):
        i = self.counters.get(group, 0)
        self.counters[group] = i + 1
        return i


PRIO = _Prio()


def install():
    d2shim.install()
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.structures import Boxes, ImageList
    m = sys.modules
    ShapeSpec = m["detectron2.layers"].ShapeSpec

    # ------------------------------------------------------------------ backbone / meta arch
    class Backbone(nn.Module):
        """d2 v0.5 modeling/backbone/backbone.py: size_divisibility defaults to 0."""

        @property
        def size_divisibility(self):
            return 0

    class GeneralizedRCNN(nn.Module):
        """d2 v0.5 modeling/meta_arch/rcnn.py (training-side surface used by pt/modeling/meta_arch/rcnn.py)."""

        def __init__(self, *, backbone, proposal_generator, roi_heads, pixel_mean, pixel_std, input_format=None,
                     vis_period=0):
            super().__init__()
            self.backbone = backbone
            self.proposal_generator = proposal_generator
            self.roi_heads = roi_heads
            self.register_buffer("pixel_mean", torch.tensor(pixel_mean).view(-1, 1, 1), False)
            self.register_buffer("pixel_std", torch.tensor(pixel_std).view(-1, 1, 1), False)

        @property
        def device(self):
            return self.pixel_mean.device

        def preprocess_image(self, batched_inputs):
            images = [x["image"].to(self.device) for x in batched_inputs]
            images = [(x - self.pixel_mean) / self.pixel_std for x in images]
            # ImageList.from_tensors with size_divisibility 0: zero-pad bottom/right to the batch maximum
            sizes = [tuple(i.shape[-2:]) for i in images]
            H, W = max(s[0] for s in sizes), max(s[1] for s in sizes)
            batch = images[0].new_zeros(len(images), images[0].shape[0], H, W)
            for k, i in enumerate(images):
                batch[k, :, :i.shape[-2], :i.shape[-1]] = i
            return ImageList(batch, sizes)

        def inference(self, batched_inputs, detected_instances=None, do_postprocess=True):
            """d2 v0.5 GeneralizedRCNN.inference (what pt/modeling/meta_arch/rcnn.py:33-34 calls in eval mode)."""
            assert not self.training and detected_instances is None
            images = self.preprocess_image(batched_inputs)
            features = self.backbone(images.tensor)
            proposals, _ = self.proposal_generator(images, features, None)
            results, _ = self.roi_heads(images, features, proposals, None)
            if do_postprocess:
                return GeneralizedRCNN._postprocess(results, batched_inputs, images.image_sizes)
            return results

        @staticmethod
        def _postprocess(instances, batched_inputs, image_sizes):
            """d2 v0.5 GeneralizedRCNN._postprocess + modeling/postprocessing.py detector_postprocess (boxes only)."""
            processed = []
            for res, inp, size in zip(instances, batched_inputs, image_sizes):
                height, width = inp.get("height", size[0]), inp.get("width", size[1])
                scale_x, scale_y = width / res.image_size[1], height / res.image_size[0]
                out = type(res)((height, width), **res.get_fields())
                boxes = out.pred_boxes
                t = boxes.tensor.clone()
                t[:, 0::2] *= scale_x   # Boxes.scale
                t[:, 1::2] *= scale_y
                boxes = Boxes(t)
                boxes.clip(out.image_size)
                out.pred_boxes = boxes
                processed.append({"instances": out[boxes.nonempty()]})
            return processed

    # ------------------------------------------------------------------ anchors
    class DefaultAnchorGenerator(nn.Module):
        """d2 v0.5 modeling/anchor_generator.py DefaultAnchorGenerator (single level)."""
        box_dim = 4

        def __init__(self, *, sizes, aspect_ratios, strides, offset=0.0):
            super().__init__()
            self.strides, self.offset = strides, offset
            cells = []
            for s, a in zip(sizes, aspect_ratios):
                anchors = []
                for size in s:
                    area = size ** 2.0
                    for ratio in a:
                        w = (area / ratio) ** 0.5
                        h = ratio * w
                        anchors.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
                cells.append(torch.tensor(anchors))
            self.cell_anchors = cells

        @property
        def num_anchors(self):
            return [len(c) for c in self.cell_anchors]

        def forward(self, features):
            out = []
            create = m["detectron2.modeling.anchor_generator"]._create_grid_offsets
            for f, stride, base in zip(features, self.strides, self.cell_anchors):
                sx, sy = create(f.shape[-2:], stride, self.offset, base.device)
                shifts = torch.stack((sx, sy, sx, sy), dim=1)
                out.append(Boxes((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4)))
            return out

    # ------------------------------------------------------------------ RPN
    class StandardRPNHead(nn.Module):
        """d2 v0.5 StandardRPNHead: 3x3 conv + ReLU, 1x1 objectness, 1x1 anchor deltas; N(0, 0.01) / 0 init."""

        def __init__(self, *, in_channels, num_anchors, box_dim=4, conv_dims=(-1,)):
            super().__init__()
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)
            self.objectness_logits = nn.Conv2d(in_channels, num_anchors, kernel_size=1, stride=1)
            self.anchor_deltas = nn.Conv2d(in_channels, num_anchors * box_dim, kernel_size=1, stride=1)
            for layer in (self.conv, self.objectness_logits, self.anchor_deltas):
                nn.init.normal_(layer.weight, std=0.01)
                nn.init.constant_(layer.bias, 0)

        def forward(self, features):
            logits, deltas = [], []
            for x in features:
                t = F.relu(self.conv(x))
                logits.append(self.objectness_logits(t))
                deltas.append(self.anchor_deltas(t))
            return logits, deltas

    class RPN(nn.Module):
        """d2 v0.5 proposal_generator/rpn.py RPN: constructor attributes, _subsample_labels, _decode_proposals."""

        def __init__(self, *, in_features, head, anchor_generator, anchor_matcher, box2box_transform,
                     batch_size_per_image, positive_fraction, pre_nms_topk, post_nms_topk, nms_thresh=0.7,
                     min_box_size=0.0, anchor_boundary_thresh=-1.0, loss_weight=1.0, box_reg_loss_type="smooth_l1",
                     smooth_l1_beta=0.0):
            super().__init__()
            self.in_features = in_features
            self.rpn_head = head
            self.anchor_generator = anchor_generator
            self.anchor_matcher = anchor_matcher
            self.box2box_transform = box2box_transform
            self.batch_size_per_image = batch_size_per_image
            self.positive_fraction = positive_fraction
            self.pre_nms_topk = {True: pre_nms_topk[0], False: pre_nms_topk[1]}
            self.post_nms_topk = {True: post_nms_topk[0], False: post_nms_topk[1]}
            self.nms_thresh = nms_thresh
            self.min_box_size = float(min_box_size)
            self.anchor_boundary_thresh = anchor_boundary_thresh
            if isinstance(loss_weight, float):
                loss_weight = {"loss_rpn_cls": loss_weight, "loss_rpn_loc": loss_weight}
            self.loss_weight = loss_weight
            self.box_reg_loss_type = box_reg_loss_type
            self.smooth_l1_beta = smooth_l1_beta

        def _subsample_labels(self, label):
            """d2 v0.5: subsample_labels(label, 256, frac, 0) then rewrite to -1 / 1 / 0 (randperm -> injected)."""
            i = PRIO.next_index("rpn")
            n = label.numel()
            pos, neg = O.subsample_labels(label, self.batch_size_per_image, self.positive_fraction, 0,
                                          PRIO.provider(("rpn_pos", i), n), PRIO.provider(("rpn_neg", i), n))
            label.fill_(-1)
            label.scatter_(0, pos, 1)
            label.scatter_(0, neg, 0)
            return label

        def _decode_proposals(self, anchors, pred_anchor_deltas):
            N = pred_anchor_deltas[0].shape[0]
            proposals = []
            for anchors_i, deltas_i in zip(anchors, pred_anchor_deltas):
                B = anchors_i.tensor.size(1)
                deltas_i = deltas_i.reshape(-1, B)
                a = anchors_i.tensor.unsqueeze(0).expand(N, -1, -1).reshape(-1, B)
                proposals.append(self.box2box_transform.apply_deltas(deltas_i, a).view(N, -1, B))
            return proposals

    # ------------------------------------------------------------------ ROI heads
    class ROIPooler(nn.Module):
        """d2 v0.5 poolers.py, single level, ROIAlignV2 = torchvision roi_align(aligned=True)."""

        def __init__(self, output_size, scales, sampling_ratio, pooler_type):
            super().__init__()
            assert len(scales) == 1 and pooler_type == "ROIAlignV2"
            self.output_size, self.scale, self.sampling_ratio = output_size, scales[0], sampling_ratio

        def forward(self, x, box_lists):
            from torchvision.ops import roi_align
            rois = torch.cat([torch.cat([torch.full((len(b), 1), float(i)), b.tensor], dim=1)
                              for i, b in enumerate(box_lists)], dim=0)
            return roi_align(x[0], rois, self.output_size, self.scale, self.sampling_ratio, True)

    class FastRCNNConvFCHead(nn.Sequential):
        """d2 v0.5 box_head.py with NUM_CONV 0, NUM_FC 2: flatten, fc1, ReLU, fc2, ReLU (c2_xavier_fill)."""

        def __init__(self, input_shape, fc_dim=1024, num_fc=2):
            super().__init__()
            size = input_shape.channels * input_shape.height * input_shape.width
            self.add_module("flatten", nn.Flatten())
            for k in range(num_fc):
                fc = nn.Linear(size, fc_dim)
                nn.init.kaiming_uniform_(fc.weight, a=1)
                nn.init.constant_(fc.bias, 0)
                self.add_module("fc{}".format(k + 1), fc)
                self.add_module("fc_relu{}".format(k + 1), nn.ReLU())
                size = fc_dim
            self._out = fc_dim

        @property
        def output_shape(self):
            return ShapeSpec(channels=self._out)

    class StandardROIHeads(nn.Module):
        """d2 v0.5 roi_heads.py ROIHeads / StandardROIHeads: attributes and _sample_proposals."""

        def __init__(self, *, box_in_features, box_pooler, box_head, box_predictor, num_classes, batch_size_per_image,
                     positive_fraction, proposal_matcher, proposal_append_gt=True, train_on_pred_boxes=False):
            super().__init__()
            self.in_features = self.box_in_features = box_in_features
            self.box_pooler, self.box_head, self.box_predictor = box_pooler, box_head, box_predictor
            self.num_classes = num_classes
            self.batch_size_per_image = batch_size_per_image
            self.positive_fraction = positive_fraction
            self.proposal_matcher = proposal_matcher
            self.proposal_append_gt = proposal_append_gt
            self.train_on_pred_boxes = train_on_pred_boxes

        def _sample_proposals(self, matched_idxs, matched_labels, gt_classes):
            has_gt = gt_classes.numel() > 0
            if has_gt:
                gt_classes = gt_classes[matched_idxs]
                gt_classes[matched_labels == 0] = self.num_classes
                gt_classes[matched_labels == -1] = -1
            else:
                gt_classes = torch.zeros_like(matched_idxs) + self.num_classes
            i = PRIO.next_index("roi")
            n = gt_classes.numel()
            fg, bg = O.subsample_labels(gt_classes, self.batch_size_per_image, self.positive_fraction, self.num_classes,
                                        PRIO.provider(("roi_pos", i), n), PRIO.provider(("roi_neg", i), n))
            sampled = torch.cat([fg, bg], dim=0)
            return sampled, gt_classes[sampled]

    class FastRCNNOutputLayers(nn.Module):
        """d2 v0.5 fast_rcnn.py FastRCNNOutputLayers: layers, forward, losses (box_reg_loss is overridden by the
        reference's subclass, fast_rcnn.py:265-336)."""

        def __init__(self, input_shape, *, box2box_transform, num_classes, test_score_thresh=0.0, test_nms_thresh=0.5,
                     test_topk_per_image=100, cls_agnostic_bbox_reg=False, smooth_l1_beta=0.0,
                     box_reg_loss_type="smooth_l1", loss_weight=1.0):
            super().__init__()
            if isinstance(input_shape, int):
                input_shape = ShapeSpec(channels=input_shape)
            self.num_classes = num_classes
            size = input_shape.channels * (input_shape.width or 1) * (input_shape.height or 1)
            self.cls_score = nn.Linear(size, num_classes + 1)
            nreg = 1 if cls_agnostic_bbox_reg else num_classes
            self.bbox_pred = nn.Linear(size, nreg * len(box2box_transform.weights))
            self.box2box_transform = box2box_transform
            self.smooth_l1_beta = smooth_l1_beta
            self.test_score_thresh = test_score_thresh
            self.test_nms_thresh = test_nms_thresh
            self.test_topk_per_image = test_topk_per_image
            self.box_reg_loss_type = box_reg_loss_type
            if isinstance(loss_weight, float):
                loss_weight = {"loss_cls": loss_weight, "loss_box_reg": loss_weight}
            self.loss_weight = loss_weight

        def forward(self, x):
            if x.dim() > 2:
                x = torch.flatten(x, start_dim=1)
            return self.cls_score(x), self.bbox_pred(x)

        def losses(self, predictions, proposals):
            scores, proposal_deltas = predictions
            cat = m["detectron2.layers"].cat
            gt_classes = cat([p.gt_classes for p in proposals], dim=0) if len(proposals) else torch.empty(0)
            if len(proposals):
                proposal_boxes = cat([p.proposal_boxes.tensor for p in proposals], dim=0)
                gt_boxes = cat([(p.gt_boxes if p.has("gt_boxes") else p.proposal_boxes).tensor for p in proposals], dim=0)
            else:
                proposal_boxes = gt_boxes = torch.empty((0, 4), device=proposal_deltas.device)
            losses = {
                "loss_cls": m["detectron2.layers"].cross_entropy(scores, gt_classes, reduction="mean"),
                "loss_box_reg": self.box_reg_loss(proposal_boxes, gt_boxes, proposal_deltas, gt_classes),
            }
            return {k: v * self.loss_weight.get(k, 1.0) for k, v in losses.items()}

    # ------------------------------------------------------------------ register
    m["detectron2.modeling.backbone.backbone"].Backbone = Backbone
    m["detectron2.modeling.backbone"].Backbone = Backbone
    m["detectron2.modeling.meta_arch.rcnn"].GeneralizedRCNN = GeneralizedRCNN
    ag = m["detectron2.modeling.anchor_generator"]
    ag.DefaultAnchorGenerator = DefaultAnchorGenerator
    for name in ("detectron2.modeling.proposal_generator", "detectron2.modeling.proposal_generator.rpn"):
        m[name].RPN, m[name].StandardRPNHead = RPN, StandardRPNHead
    m["detectron2.modeling.roi_heads"].StandardROIHeads = StandardROIHeads
    m["detectron2.modeling.poolers"].ROIPooler = ROIPooler
    m["detectron2.modeling.roi_heads.box_head"].FastRCNNConvFCHead = FastRCNNConvFCHead
    m["detectron2.modeling.roi_heads.fast_rcnn"].FastRCNNOutputLayers = FastRCNNOutputLayers
    return dict(Backbone=Backbone, GeneralizedRCNN=GeneralizedRCNN, DefaultAnchorGenerator=DefaultAnchorGenerator,
                RPN=RPN, StandardRPNHead=StandardRPNHead, StandardROIHeads=StandardROIHeads, ROIPooler=ROIPooler,
                FastRCNNConvFCHead=FastRCNNConvFCHead, FastRCNNOutputLayers=FastRCNNOutputLayers, ShapeSpec=ShapeSpec)
