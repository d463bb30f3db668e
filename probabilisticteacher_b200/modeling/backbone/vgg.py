"""VGG16 backbone on the B200 path: the counterpart of `pt/modeling/backbone/vgg.py:36-72` (VGGBlock:
3x3 conv + bias + ReLU, 2x2 max pool after blocks 1-4), `:94-165` (VGG, stride-16 `vgg_block5` output) and
`:189-230` (build_vgg_backbone, FREEZE_AT). Every conv is one launch of the tcgen05 implicit-GEMM
kernel over fp16 NHWC-flat activations; backward is explicit (data- and weight-gradient GEMMs)."""
import torch
from torch import nn

from ... import ops
from ..._lib import call
from ..registry import BACKBONE_REGISTRY


class VGG(nn.Module):
    def __init__(self, arena, loss_scale):
        super().__init__()
        self.arena = arena
        self.loss_scale = loss_scale
        self._out_features = ["vgg_block5"]
        self._out_feature_strides = {"vgg_block5": 16}
        self._out_feature_channels = {"vgg_block5": arena.C}

    def output_shape(self):
        return {"vgg_block5": dict(channels=self.arena.C, stride=16)}

    def forward(self, x: ops.FlatAct, save=False):
        """x: output of vgg_block1.conv1 (+ReLU), produced from the uint8 images by the fused
        pre-processing + first-conv kernel (see GuassianGeneralizedRCNN.preprocess_image).
        Returns ({"vgg_block5": FlatAct}, records) where records hold what backward needs."""
        ar = self.arena
        specs = ar.conv_specs
        records = []
        block = 1
        if ar.precision == "f16x3":
            # split-fp16 fp32-equivalent precision: x and every activation are [hi | lo | hi] triples
            for name, cin, cout, trainable in specs[1:]:
                b = int(name.split("vgg_block")[1][0])
                pooled = False
                if b != block:
                    x = ops.maxpool2x2_x3(x)
                    block = b
                    pooled = True
                w3, alpha = ar.x3view(name + ".weight")
                y = ops.conv3x3_x3(x, w3, alpha, ar.view(name + ".bias"))
                if save and trainable:
                    records.append(dict(name=name, x=x, y=y, pooled=pooled, cin=cin, cout=cout))
                x = y
            return {"vgg_block5": x}, records
        for name, cin, cout, trainable in specs[1:]:
            b = int(name.split("vgg_block")[1][0])
            pooled = False
            if b != block:
                x = ops.maxpool2x2(x)
                block = b
                pooled = True
            w = ar.hview(name + ".weight").view(cout, 9 * cin)
            y = ops.conv3x3(x, w, ar.view(name + ".bias"), relu=True)
            if save and trainable:
                records.append(dict(name=name, x=x, y=y, pooled=pooled, cin=cin, cout=cout))
            x = y
        return {"vgg_block5": x}, records

    def backward(self, records, dz: ops.FlatAct):
        """dz: gradient w.r.t. the pre-ReLU output of the last conv (already ReLU-masked), scaled by
        the loss scale. Accumulates weight / bias gradients into the arena."""
        ar = self.arena
        inv = 1.0 / self.loss_scale
        if ar.precision == "f16x3":
            return self._backward_x3(records, dz)
        for i in range(len(records) - 1, -1, -1):
            r = records[i]
            gw = ar.gview(r["name"] + ".weight").view(r["cout"], 9 * r["cin"])
            ops.conv3x3_wgrad(dz, r["x"], gw, scale=inv, bias_out=ar.gview(r["name"] + ".bias"))
            if i == 0:
                break
            wd = ar.dgrad_half[r["name"]]
            if r["pooled"]:
                dp = ops.conv3x3(dz, wd, None, relu=False)
                dz = ops.maxpool2x2_relu_bwd(records[i - 1]["y"], dp)
            else:
                dz = ops.conv3x3(dz, wd, None, aux=r["x"].t)


    def _backward_x3(self, records, dz: ops.FlatAct):
        """Same chain in the f16x3 precision (the reference's fp32 autograd, pt/engine/trainer.py:383-386): dz and
        every saved activation are triples; data gradients that feed a max-pool backward are stored in fp32."""
        ar = self.arena
        inv = 1.0 / self.loss_scale
        for i in range(len(records) - 1, -1, -1):
            r = records[i]
            gw = ar.gview(r["name"] + ".weight").view(r["cout"], 9 * r["cin"])
            ops.conv3x3_wgrad_x3(dz, r["x"], gw, r["cout"], r["cin"], scale=inv,
                                 bias_out=ar.gview(r["name"] + ".bias"))
            if i == 0:
                break
            wd3, alpha = ar.dgrad_x3[r["name"]]
            if r["pooled"]:
                dp = ops.conv3x3_dgrad_x3(dz, wd3, alpha)
                dz = ops.maxpool2x2_relu_bwd_x3(records[i - 1]["y"], dp)
            else:
                dz = ops.conv3x3_dgrad_x3(dz, wd3, alpha, aux=r["x"].t)


@BACKBONE_REGISTRY.register()
def build_vgg_backbone(cfg, arena, loss_scale):
    assert cfg.MODEL.VGG.DEPTH == 16, "only VGG16 is on the hot path (configs/Guassian-RCNN-VGG.yaml:8)"
    if cfg.MODEL.BACKBONE.FREEZE_AT < 1:
        # vgg.py:175-180 freezes blocks 1..FREEZE_AT (detectron2 default 2, which every reference config keeps). The
        # first conv is fused with the uint8 pre-processing and has no weight-gradient kernel: refusing is better
        # than training with a conv1_1 that silently never receives a gradient.
        raise ValueError("MODEL.BACKBONE.FREEZE_AT must be >= 1 on this path (vgg_block1.conv1 is forward-only)")
    return VGG(arena, loss_scale)
