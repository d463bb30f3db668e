"""Anchor generators: `pt/modeling/anchor_generator.py:31-164` (DifferentiableAnchorGenerator: 9 learnable
(w, h) pairs) and detectron2's DefaultAnchorGenerator (base yaml, sizes x aspect ratios)."""
import math

import torch
from torch import nn

from .._lib import call
from .registry import ANCHOR_GENERATOR_REGISTRY


class _GridMixin:
    def grid(self, cell, H, W):
        R = H * W * self.num_cell
        out = torch.empty(R, 4, dtype=torch.float32, device=cell.device)
        call("ptb200_anchor_grid", cell, self.num_cell, H, W, float(self.stride), float(self.offset), out)
        return out


@ANCHOR_GENERATOR_REGISTRY.register()
class DifferentiableAnchorGenerator(nn.Module, _GridMixin):
    box_dim = 4

    def __init__(self, cfg, arena, stride=16):
        super().__init__()
        self.arena = arena
        self.stride = stride
        self.offset = cfg.MODEL.ANCHOR_GENERATOR.OFFSET
        self.num_cell = arena.A
        self.differentiable = True
        # the learnable (w, h) pairs start from cfg.MODEL.ANCHOR_GENERATOR.ANCHOR (anchor_generator.py:66-72), so a
        # model built from cfg alone (e.g. followed by a backbone-only weight file) has usable anchors
        anchor = torch.tensor(cfg.MODEL.ANCHOR_GENERATOR.ANCHOR[0], dtype=torch.float32)
        with torch.no_grad():
            arena.view("proposal_generator.anchor_generator.anchor_0").copy_(anchor.to(arena.device))

    def forward(self, H, W):
        wh = self.arena.view("proposal_generator.anchor_generator.anchor_0")
        cell = torch.empty(self.num_cell, 4, dtype=torch.float32, device=wh.device)
        call("ptb200_cell_anchors_from_wh", wh, self.num_cell, cell)
        return self.grid(cell, H, W)


@ANCHOR_GENERATOR_REGISTRY.register()
class DefaultAnchorGenerator(nn.Module, _GridMixin):
    box_dim = 4

    def __init__(self, cfg, arena, stride=16):
        super().__init__()
        sizes = cfg.MODEL.ANCHOR_GENERATOR.SIZES[0]
        ratios = cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS[0]
        cells = []
        for s in sizes:
            area = s ** 2.0
            for r in ratios:
                w = math.sqrt(area / r)
                h = r * w
                cells.append([-w / 2.0, -h / 2.0, w / 2.0, h / 2.0])
        self.register_buffer("cell", torch.tensor(cells, dtype=torch.float32, device=arena.device), persistent=False)
        self.stride = stride
        self.offset = cfg.MODEL.ANCHOR_GENERATOR.OFFSET
        self.num_cell = len(cells)
        self.differentiable = False

    def forward(self, H, W):
        return self.grid(self.cell, H, W)
