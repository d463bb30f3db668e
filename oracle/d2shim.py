"""Stub `detectron2` / `fvcore` packages -- TEST INFRASTRUCTURE, only used by oracle/make_golden.py.

The reference's hot-path files import detectron2==0.5 and fvcore, which are neither vendored in the
reference tree nor installable here. `install()` registers just enough fake modules in `sys.modules`
for `/root/reference/pt/modeling/*.py` to import UNMODIFIED, so that their own functions (losses,
proposal selection, pseudo-label filter, labelling) can be executed to freeze golden vectors. The
detectron2 behaviours the shim has to supply (Boxes, Instances, pairwise_iou, Matcher,
subsample_labels, batched_nms) are restated from the v0.5 release; everything that is only needed as
a base class or decorator is an empty placeholder.
"""
import sys
import types

import torch


def _mod(name):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


def install():
    if "detectron2" in sys.modules and getattr(sys.modules["detectron2"], "_ptb200_shim", False):
        return
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from probabilisticteacher_b200.structures import Boxes, ImageList, Instances
    from oracle import pt_oracle as O

    d2 = _mod("detectron2")
    d2._ptb200_shim = True

    # ---- config
    cfgm = _mod("detectron2.config")

    class CfgNode(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    def configurable(init_func=None, *, from_config=None):
        if init_func is not None:
            return init_func
        return lambda f: f
    cfgm.CfgNode = CfgNode
    cfgm.configurable = configurable

    # ---- layers
    lay = _mod("detectron2.layers")

    def cat(tensors, dim=0):
        if len(tensors) == 1:
            return tensors[0]
        return torch.cat(tensors, dim)

    def batched_nms(boxes, scores, idxs, iou_threshold):
        """d2 v0.5 layers/nms.py -> torchvision batched_nms; per-class NMS on un-shifted boxes
        (torchvision `_batched_nms_vanilla`), through torchvision's own CPU nms kernel."""
        from torchvision.ops import nms as tv_nms
        boxes = boxes.float()
        keep_mask = torch.zeros_like(scores, dtype=torch.bool)
        for c in torch.unique(idxs):
            ci = torch.where(idxs == c)[0]
            keep_mask[ci[tv_nms(boxes[ci], scores[ci], iou_threshold)]] = True
        keep = torch.where(keep_mask)[0]
        return keep[scores[keep].argsort(descending=True, stable=True)]

    class ShapeSpec:
        def __init__(self, channels=None, height=None, width=None, stride=None):
            self.channels, self.height, self.width, self.stride = channels, height, width, stride

    def cross_entropy(input, target, *, reduction="mean", **kw):
        if target.numel() == 0 and reduction == "mean":
            return input.sum() * 0.0
        return torch.nn.functional.cross_entropy(input, target, reduction=reduction, **kw)

    def nonzero_tuple(x):
        return x.nonzero().unbind(1)

    class Conv2d(torch.nn.Conv2d):
        def __init__(self, *args, **kwargs):
            self.norm = kwargs.pop("norm", None)
            self.activation = kwargs.pop("activation", None)
            super().__init__(*args, **kwargs)

    class CNNBlockBase(torch.nn.Module):
        def __init__(self, in_channels, out_channels, stride):
            super().__init__()
            self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            return self

    lay.cat, lay.batched_nms, lay.ShapeSpec = cat, batched_nms, ShapeSpec
    lay.cross_entropy, lay.nonzero_tuple = cross_entropy, nonzero_tuple
    lay.Conv2d, lay.CNNBlockBase, lay.get_norm = Conv2d, CNNBlockBase, (lambda norm, ch: None)

    # ---- structures
    st = _mod("detectron2.structures")
    st.Boxes, st.Instances, st.ImageList = Boxes, Instances, ImageList
    st.RotatedBoxes = type("RotatedBoxes", (), {})
    st.pairwise_iou = lambda b1, b2: O.pairwise_iou(b1.tensor, b2.tensor)

    # ---- utils
    _mod("detectron2.utils")
    ev = _mod("detectron2.utils.events")

    class _Storage:
        def put_scalar(self, *a, **k):
            pass
    ev.get_event_storage = lambda: _Storage()
    mem = _mod("detectron2.utils.memory")
    mem.retry_if_cuda_oom = lambda f: f
    reg = _mod("detectron2.utils.registry")

    class Registry(dict):
        def __init__(self, name=""):
            super().__init__()

        def register(self, obj=None):
            def deco(o):
                self[o.__name__] = o
                return o
            return deco(obj) if obj is not None else deco
    reg.Registry = Registry

    # ---- modeling
    _mod("detectron2.modeling")
    ag = _mod("detectron2.modeling.anchor_generator")
    ag.ANCHOR_GENERATOR_REGISTRY = Registry()
    ag.build_anchor_generator = lambda cfg, shape: None

    def _broadcast_params(params, num_features, name):
        if not isinstance(params[0], (list, tuple)):
            return [params] * num_features
        if len(params) == 1:
            return list(params) * num_features
        return params

    def _create_grid_offsets(size, stride, offset, device):
        gh, gw = size
        sx = torch.arange(offset * stride, gw * stride, step=stride, dtype=torch.float32, device=device)
        sy = torch.arange(offset * stride, gh * stride, step=stride, dtype=torch.float32, device=device)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        return xx.reshape(-1), yy.reshape(-1)
    ag._broadcast_params, ag._create_grid_offsets = _broadcast_params, _create_grid_offsets

    mt = _mod("detectron2.modeling.matcher")

    class Matcher:
        def __init__(self, thresholds, labels, allow_low_quality_matches=False):
            self.thresholds, self.labels, self.allow = list(thresholds), list(labels), allow_low_quality_matches

        def __call__(self, m):
            return O.matcher(m, self.thresholds, self.labels, self.allow)
    mt.Matcher = Matcher

    pg = _mod("detectron2.modeling.proposal_generator")
    pgb = _mod("detectron2.modeling.proposal_generator.build")
    pgr = _mod("detectron2.modeling.proposal_generator.rpn")
    pgu = _mod("detectron2.modeling.proposal_generator.proposal_utils")
    pgb.PROPOSAL_GENERATOR_REGISTRY = Registry()
    pgr.RPN_HEAD_REGISTRY = Registry()
    pgr.build_rpn_head = lambda cfg, shape: None
    pgu._is_tracing = lambda: False

    class RPN(torch.nn.Module):
        pass

    class StandardRPNHead(torch.nn.Module):
        pass
    pg.RPN, pg.StandardRPNHead = RPN, StandardRPNHead
    pgr.RPN, pgr.StandardRPNHead = RPN, StandardRPNHead

    rh = _mod("detectron2.modeling.roi_heads")
    rh.ROI_HEADS_REGISTRY = Registry()
    rh.StandardROIHeads = type("StandardROIHeads", (torch.nn.Module,), {})
    bh = _mod("detectron2.modeling.roi_heads.box_head")
    bh.build_box_head = lambda cfg, shape: None
    fr = _mod("detectron2.modeling.roi_heads.fast_rcnn")
    fr.FastRCNNOutputLayers = type("FastRCNNOutputLayers", (torch.nn.Module,), {})
    po = _mod("detectron2.modeling.poolers")
    po.ROIPooler = type("ROIPooler", (torch.nn.Module,), {})
    _mod("detectron2.modeling.backbone")
    bb = _mod("detectron2.modeling.backbone.backbone")
    bb.Backbone = type("Backbone", (torch.nn.Module,), {})
    bbb = _mod("detectron2.modeling.backbone.build")
    bbb.BACKBONE_REGISTRY = Registry()
    _mod("detectron2.modeling.meta_arch")
    mab = _mod("detectron2.modeling.meta_arch.build")
    mab.META_ARCH_REGISTRY = Registry()
    mar = _mod("detectron2.modeling.meta_arch.rcnn")
    mar.GeneralizedRCNN = type("GeneralizedRCNN", (torch.nn.Module,), {})

    # ---- fvcore
    _mod("fvcore")
    fn = _mod("fvcore.nn")

    def smooth_l1_loss(input, target, beta, reduction="none"):
        n = torch.abs(input - target)
        loss = n if beta < 1e-5 else torch.where(n < beta, 0.5 * n ** 2 / beta, n - 0.5 * beta)
        return loss.sum() if reduction == "sum" else (loss.mean() if reduction == "mean" else loss)
    fn.smooth_l1_loss = smooth_l1_loss
    fn.giou_loss = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("giou not on the hot path"))
    wi = _mod("fvcore.nn.weight_init")

    def c2_msra_fill(module):
        torch.nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
        if module.bias is not None:
            torch.nn.init.constant_(module.bias, 0)
    wi.c2_msra_fill = c2_msra_fill


def subsample_labels_with_prio(labels, num_samples, positive_fraction, bg_label, prio_pos, prio_neg):
    """d2 v0.5 sampling.subsample_labels with torch.randperm replaced by the injected-priority
    permutation (see oracle/pt_oracle.py:_perm_from_prio)."""
    from oracle import pt_oracle as O
    return O.subsample_labels(labels, num_samples, positive_fraction, bg_label, prio_pos, prio_neg)
