"""Flat parameter arena of one detector (student or teacher).

All parameters live in ONE contiguous fp32 buffer laid out for the kernels (HBM-resident, 16-byte
aligned segments), so that the EMA teacher update, the gradient all-reduce, the grad-norm clip and
the SGD step are single streaming passes (`pt/engine/trainer.py:431-449,592-603,386`), and the fp16
GEMM operands are produced by one cast pass over the same layout.

Internal layouts (reference layout in brackets; `state_dict()` converts):
  conv weight   [Cout][ky][kx][Cin]           ([Cout][Cin][3][3])
  fc1 weight    [1024][ph*7+pw][512]          ([1024][512*49], channel-major flatten of NCHW)
  RPN 1x1 heads one block [128][512]: rows 0..8 objectness, 9..80 anchor deltas, rest zero padding
  predictor     one block [128][1024]: rows 0..K cls_score, K+1..K+8K bbox_pred, rest zero padding
Parameter names are the reference's state_dict keys (checkpoint row).
"""
import math
from collections import OrderedDict

import torch

from ._lib import call

VGG16 = [[64, 64], [128, 128], [256, 256, 256], [512, 512, 512], [512, 512, 512]]


def _round64(n):
    """Segments start on 64-element boundaries: 128 B in the fp16 operand arena, so that every 128-byte TMA box
    row of a weight tile is one aligned L2 line. (With 16-byte alignment the fc1 weight block sat 48 B off a line:
    each box row touched two lines and the L2-bound fc1 GEMM ran at 270 instead of 160 us.)"""
    return (n + 63) // 64 * 64


class Segment:
    __slots__ = ("name", "shape", "offset", "numel", "trainable", "kind")

    def __init__(self, name, shape, offset, trainable, kind):
        self.name, self.shape, self.offset, self.trainable, self.kind = name, tuple(shape), offset, trainable, kind
        self.numel = int(math.prod(shape))


class ParamArena:
    def __init__(self, num_classes=8, num_cell=9, fc_dim=1024, pooled=7, freeze_at=2,
                 differentiable_anchors=True, device="cuda", with_grads=True):
        self.K = num_classes
        self.A = num_cell
        self.fc_dim = fc_dim
        self.pooled = pooled
        self.device = torch.device(device)
        segs = []
        off = 0

        def add(name, shape, trainable, kind):
            nonlocal off
            s = Segment(name, shape, off, trainable, kind)
            segs.append(s)
            off += _round64(s.numel)
            return s

        convs = []
        cin = 3
        for bi, chans in enumerate(VGG16, start=1):
            for ci, cout in enumerate(chans, start=1):
                convs.append((f"backbone.vgg_block{bi}.0.conv{ci}", cin, cout, bi > freeze_at))
                cin = cout
        self.conv_specs = convs
        C = cin
        for tr in (False, True):  # frozen segments first so that the trainable part is a contiguous suffix
            for name, ci_, co_, trainable in convs:
                if trainable == tr:
                    add(name + ".weight", (co_, 3, 3, ci_), trainable, "conv")
                    add(name + ".bias", (co_,), trainable, "vec")
        add("proposal_generator.rpn_head.conv.weight", (C, 3, 3, C), True, "conv")
        add("proposal_generator.rpn_head.conv.bias", (C,), True, "vec")
        add("proposal_generator.rpn_head._heads.weight", (128, C), True, "rpn_heads_w")
        add("proposal_generator.rpn_head._heads.bias", (128,), True, "rpn_heads_b")
        if differentiable_anchors:
            add("proposal_generator.anchor_generator.anchor_0", (num_cell, 2), True, "mat")
        fin = pooled * pooled
        add("roi_heads.box_head.fc1.weight", (fc_dim, fin, C), True, "fc1")
        add("roi_heads.box_head.fc1.bias", (fc_dim,), True, "vec")
        add("roi_heads.box_head.fc2.weight", (fc_dim, fc_dim), True, "mat")
        add("roi_heads.box_head.fc2.bias", (fc_dim,), True, "vec")
        add("roi_heads.box_predictor._heads.weight", (128, fc_dim), True, "pred_w")
        add("roi_heads.box_predictor._heads.bias", (128,), True, "pred_b")
        self.segments = OrderedDict((s.name, s) for s in segs)
        self.total = off
        self.trainable_start = min(s.offset for s in segs if s.trainable)
        self.C = C
        self.data = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        self.half = torch.zeros(self.total, dtype=torch.float16, device=self.device)
        self.conv1_half = torch.zeros(64, 32, dtype=torch.float16, device=self.device)
        self.precision = "f16"  # "f16x3": split-fp16 fp32-equivalent precision, see pack_x3
        self.x3 = None
        self._x3_exp = None
        self.dgrad_x3 = None
        # pixel statistics of the model (set by GuassianGeneralizedRCNN): the f16x3 first conv folds them into its
        # weight / bias operands
        self.pixel_mean = (103.530, 116.280, 123.675)
        self.pixel_std = (1.0, 1.0, 1.0)
        self.conv1_x3 = None  # (wpack3 fp16 [64][64], bias table fp32 [10][64], alpha)
        if with_grads:
            n = self.total - self.trainable_start
            # the gradient arena is followed by a small fp32 metrics tail: the 8 loss scalars of the step ride on the
            # gradient all-reduce (the reference averages its logged metrics over ranks, pt/engine/trainer.py:394-429)
            self.METRICS_TAIL = 16
            self.grads_ext = torch.zeros(n + self.METRICS_TAIL, dtype=torch.float32, device=self.device)
            self.grads = self.grads_ext[:n]
            self.metrics_tail = self.grads_ext[n:]
            self.momentum = torch.zeros(n, dtype=torch.float32, device=self.device)
            self.dgrad_half = {}
        else:
            self.grads = self.momentum = None
            self.dgrad_half = None

    # ------------------------------------------------------------------ views
    def view(self, name):
        s = self.segments[name]
        return self.data[s.offset:s.offset + s.numel].view(s.shape)

    def hview(self, name):
        s = self.segments[name]
        return self.half[s.offset:s.offset + s.numel].view(s.shape)

    def head_bucket_start(self):
        """Offset (in elements of `grads`) where the parameters behind the VGG backbone start: the cut between the two
        gradient all-reduce buckets of the graph step (engine/trainer.py). Everything from here on -- RPN head, box
        head, predictor, learnable anchors -- has its final gradient as soon as both student passes have finished their
        head backward; everything before it is backbone."""
        return self.segments["proposal_generator.rpn_head.conv.weight"].offset - self.trainable_start

    def gview(self, name, flat=None):
        """View of a segment inside a flat buffer laid out like the trainable suffix (default: the gradient arena;
        the momentum arena has the same layout)."""
        s = self.segments[name]
        o = s.offset - self.trainable_start
        flat = self.grads if flat is None else flat
        return flat[o:o + s.numel].view(s.shape)

    def exposed_parameters(self, flat=None):
        """(reference name, data view, grad view or None, trainable) for every reference parameter. `flat`
        replaces the gradient arena as the buffer the third element views (used for the momentum arena)."""
        K, A = self.K, self.A
        out = []
        flat = self.grads if flat is None else flat
        for s in self.segments.values():
            v = self.view(s.name)
            g = self.gview(s.name, flat) if (s.trainable and flat is not None) else None
            if s.kind == "rpn_heads_w":
                base = "proposal_generator.rpn_head."
                out.append((base + "objectness_logits.weight", v[:A], None if g is None else g[:A], True))
                out.append((base + "anchor_deltas.weight", v[A:A * 9], None if g is None else g[A:A * 9], True))
            elif s.kind == "rpn_heads_b":
                base = "proposal_generator.rpn_head."
                out.append((base + "objectness_logits.bias", v[:A], None if g is None else g[:A], True))
                out.append((base + "anchor_deltas.bias", v[A:A * 9], None if g is None else g[A:A * 9], True))
            elif s.kind == "pred_w":
                base = "roi_heads.box_predictor."
                out.append((base + "cls_score.weight", v[:K + 1], None if g is None else g[:K + 1], True))
                out.append((base + "bbox_pred.weight", v[K + 1:9 * K + 1], None if g is None else g[K + 1:9 * K + 1], True))
            elif s.kind == "pred_b":
                base = "roi_heads.box_predictor."
                out.append((base + "cls_score.bias", v[:K + 1], None if g is None else g[:K + 1], True))
                out.append((base + "bbox_pred.bias", v[K + 1:9 * K + 1], None if g is None else g[K + 1:9 * K + 1], True))
            else:
                out.append((s.name, v, g, s.trainable))
        return out

    # ------------------------------------------------------------------ reference layout <-> internal
    @staticmethod
    def _to_ref(kind, t, C, pooled):
        if kind == "conv":
            return t.permute(0, 3, 1, 2).contiguous()
        if kind == "fc1":
            return t.permute(0, 2, 1).reshape(t.shape[0], -1).contiguous()
        return t.clone()

    def state_dict(self):
        sd = OrderedDict()
        kinds = {}
        for s in self.segments.values():
            kinds[s.name] = s.kind
        for name, v, _, _ in self.exposed_parameters():
            kind = kinds.get(name, "mat")
            t = self._to_ref(kind, v, self.C, self.pooled)
            if name.endswith("objectness_logits.weight") or name.endswith("anchor_deltas.weight"):
                t = t.reshape(t.shape[0], t.shape[1], 1, 1)
            sd[name] = t
        return sd

    def _from_ref(self, kind, src, like):
        if kind == "conv":
            src = src.permute(0, 2, 3, 1)
        elif kind == "fc1":
            src = src.reshape(src.shape[0], self.C, self.pooled * self.pooled).permute(0, 2, 1)
        return src.reshape(like.shape)

    def momentum_state_dict(self):
        """SGD momentum buffers under the reference's parameter names / layouts (trainable parameters only)."""
        kinds = {s.name: s.kind for s in self.segments.values()}
        sd = OrderedDict()
        for name, _, m, trainable in self.exposed_parameters(self.momentum):
            if m is None or not trainable:
                continue
            t = self._to_ref(kinds.get(name, "mat"), m, self.C, self.pooled)
            if name.endswith("objectness_logits.weight") or name.endswith("anchor_deltas.weight"):
                t = t.reshape(t.shape[0], t.shape[1], 1, 1)
            sd[name] = t
        return sd

    def load_momentum_state_dict(self, sd):
        """Inverse of momentum_state_dict; parameters absent from `sd` keep a zero buffer (torch SGD creates the
        buffer lazily on the first step, so a checkpoint written before step 1 has none)."""
        kinds = {s.name: s.kind for s in self.segments.values()}
        with torch.no_grad():
            self.momentum.zero_()
            for name, _, m, trainable in self.exposed_parameters(self.momentum):
                if m is None or not trainable or name not in sd:
                    continue
                src = sd[name].to(self.device, torch.float32)
                m.copy_(self._from_ref(kinds.get(name, "mat"), src, m))

    def load_state_dict(self, sd, strict=True):
        """Copies a reference-layout state dict into the arena. strict: a missing key raises KeyError (as before);
        otherwise missing keys are skipped. Returns (missing_keys, unexpected_keys)."""
        kinds = {s.name: s.kind for s in self.segments.values()}
        missing = []
        known = set()
        with torch.no_grad():
            for name, v, _, _ in self.exposed_parameters():
                known.add(name)
                if name not in sd:
                    if strict:
                        raise KeyError(f"{name} missing from state_dict")
                    missing.append(name)
                    continue
                src = sd[name].to(self.device, torch.float32)
                v.copy_(self._from_ref(kinds.get(name, "mat"), src, v))
        self._x3_exp = None  # new weights: new f16x3 scales
        self._conv1_dirty = True
        self.pack()
        return missing, [k for k in sd if k not in known]

    # ------------------------------------------------------------------ fp16 operands
    def pack(self, dgrad=None):
        """fp32 masters -> fp16 GEMM operands: one cast pass for the forward operands (same offsets),
        the K-padded first conv, and (student only) the transposed operands of the data-gradient GEMMs. In the
        f16x3 precision the weight triples (forward, and data-gradient for the student) are refreshed as well, into
        persistent buffers and with the power-of-two scales fixed at the last `refresh_x3_scales()` (no host sync:
        the call is CUDA-graph capturable once the scales exist)."""
        if dgrad is None:
            dgrad = self.dgrad_half is not None
        if self.precision == "f16x3":
            self.pack_x3(dgrad=dgrad)
            return
        call("ptb200_cast_f32_f16", self.data, self.half, self.total)
        w1 = self.view("backbone.vgg_block1.0.conv1.weight").view(64, 27)
        call("ptb200_cast_pad_rows_f16", w1, self.conv1_half, 64, 27, 32)
        if dgrad:
            self._pack_dgrad()

    # ------------------------------------------------------------------ f16x3 operands
    _X3_KINDS = ("conv", "fc1", "mat", "rpn_heads_w", "pred_w")

    def _x3_segments(self):
        # (the 3 -> 64 first conv runs in fp32 from the master weights; anchors are not GEMM operands)
        return [s for s in self.segments.values() if s.kind in self._X3_KINDS and s.shape[-1] % 64 == 0]

    def refresh_x3_scales(self):
        """Per-tensor power-of-two scale 2^e with max|W| * 2^e in [2^12, 2^13): the lo halves stay normal fp16
        numbers and the hi halves keep an 8x margin to the fp16 maximum while training moves the weights. ONE host
        sync (all maxima are reduced on the device first); called on the first pack and after load_state_dict."""
        segs = self._x3_segments()
        w1 = self.view(self.conv_specs[0][0] + ".weight")  # first conv: w / std (see _pack_conv1_x3)
        std = self._pixel_stats()[1].float()
        m = torch.stack([self.view(s.name).abs().max() for s in segs] + [(w1 / std).abs().max()]).tolist()
        self._x3_exp = {}
        for name, mx in zip([s.name for s in segs] + ["__conv1__"], m):
            e = math.floor(math.log2(8192.0 / mx)) if mx > 0 and math.isfinite(mx) else 0
            self._x3_exp[name] = max(-24, min(24, e))

    def _pixel_stats(self):
        """(mean, std) as fp64 device tensors, cached: no host -> device copy once they exist (CUDA-graph capture)."""
        key = (tuple(self.pixel_mean), tuple(self.pixel_std))
        c = getattr(self, "_pix_cache", None)
        if c is None or c[0] != key:
            c = (key, torch.tensor(self.pixel_mean, dtype=torch.float64, device=self.device),
                 torch.tensor(self.pixel_std, dtype=torch.float64, device=self.device))
            self._pix_cache = c
        return c[1], c[2]

    def _pack_conv1_x3(self):
        """Operands of ptb200_conv1_u8_f16x3_tc: raw pixels are exact in fp16, so only the weights are split --
        rows [Wh(27) 0(5) | Wl(27) 0(5)] of w / std * 2^e -- and the mean goes into a bias table: rows 0..8 hold
        S_t[co] = sum_c w[co][t][c] / std_c * mean_c, row 9 = bias - sum_t S_t (evaluated in fp64). A handful of
        small device ops, no host sync."""
        name = self.conv_specs[0][0]
        e = self._x3_exp["__conv1__"]
        w = self.view(name + ".weight").view(64, 9, 3).double()
        mean, std = self._pixel_stats()
        wp = w / std.view(1, 1, 3)
        ws = (wp * (2.0 ** e)).float().view(64, 27)
        wh = ws.half()
        wl = (ws - wh.float()).half()
        if self.conv1_x3 is None:
            self.conv1_x3 = (torch.zeros(64, 64, dtype=torch.float16, device=self.device),
                             torch.zeros(10, 64, dtype=torch.float32, device=self.device), 0.0)
        pack, table, _ = self.conv1_x3
        pack[:, :27] = wh
        pack[:, 32:59] = wl
        S = (wp * mean.view(1, 1, 3)).sum(-1)  # [64][9]
        table[:9] = S.t().float()
        table[9] = (self.view(name + ".bias").double() - S.sum(1)).float()
        self.conv1_x3 = (pack, table, 2.0 ** (-e))

    # data-gradient operands: arena segment -> (key, rows = Cout, cols = Cin, taps, flip)
    def _x3_dgrad_specs(self):
        specs = []
        first_trainable = True
        for name, cin, cout, trainable in self.conv_specs:
            if not trainable:
                continue
            if first_trainable:  # its input is the frozen part: no data-gradient needed
                first_trainable = False
                continue
            specs.append((name + ".weight", name, cout, cin, 9, 1))
        C, fc = self.C, self.fc_dim
        fin = self.pooled * self.pooled * C
        specs.append(("proposal_generator.rpn_head.conv.weight", "rpn_conv", C, C, 9, 1))
        specs.append(("proposal_generator.rpn_head._heads.weight", "rpn_heads", 128, C, 1, 0))
        specs.append(("roi_heads.box_head.fc1.weight", "fc1", fc, fin, 1, 0))
        specs.append(("roi_heads.box_head.fc2.weight", "fc2", fc, fc, 1, 0))
        specs.append(("roi_heads.box_predictor._heads.weight", "pred", 128, fc, 1, 0))
        return specs

    def pack_x3(self, dgrad=False):
        """fp32 masters -> f16x3 weight triples [Wh | Wh | Wl] of W * 2^s per tensor; caches
        {name: (triples [rows, taps*3k], alpha = 2^-s)} in `self.x3` and, with dgrad, the transposed
        data-gradient operands {key: (triples [Cin, taps*3*Cout], alpha)} in `self.dgrad_x3`."""
        if getattr(self, "_x3_exp", None) is None:
            self.refresh_x3_scales()
        if self.x3 is None:
            self.x3 = {}
        # the first conv is frozen (FREEZE_AT >= 1): the student's operands only change when weights are loaded; the
        # teacher's follow the EMA (they differ from the student's until the first copy step)
        if self.conv1_x3 is None or self.grads is None or getattr(self, "_conv1_dirty", True):
            self._pack_conv1_x3()
            self._conv1_dirty = False
        for s in self._x3_segments():
            # reduction length of one weight row per tap: Cin for convs, the whole row for fully connected layers
            # (fc1: 49 * 512, matching the [hi | lo | hi] roi rows ptb200_roi_align_fwd_f16x3 writes)
            k = s.numel // s.shape[0] if s.kind == "fc1" else s.shape[-1]
            e = self._x3_exp[s.name]
            rows = s.numel // k
            ent = self.x3.get(s.name)
            if ent is None:
                ent = (torch.empty(s.shape[0], (rows // s.shape[0]) * 3 * k, dtype=torch.float16, device=self.device),
                       2.0 ** (-e))
            self.x3[s.name] = (ent[0], 2.0 ** (-e))
            call("ptb200_split3_pack_f16", self.view(s.name), ent[0], rows, k, 2.0 ** e, 1)
        if dgrad:
            if getattr(self, "dgrad_x3", None) is None:
                self.dgrad_x3 = {}
            for seg_name, key, rows, cols, taps, flip in self._x3_dgrad_specs():
                e = self._x3_exp[seg_name]
                ent = self.dgrad_x3.get(key)
                if ent is None:
                    ent = (torch.empty(cols, taps * 3 * rows, dtype=torch.float16, device=self.device), 0.0)
                self.dgrad_x3[key] = (ent[0], 2.0 ** (-e))
                call("ptb200_transpose_pack_f16x3", self.view(seg_name), ent[0], rows, cols, taps, flip, 2.0 ** e)
        return self.x3

    def x3view(self, name):
        if self.x3 is None:
            self.pack_x3(dgrad=self.dgrad_half is not None)
        return self.x3[name]

    def _dg(self, name, shape):
        t = self.dgrad_half.get(name)
        if t is None:
            t = torch.zeros(shape, dtype=torch.float16, device=self.device)
            self.dgrad_half[name] = t
        return t

    def _pack_dgrad(self):
        first_trainable = True
        for name, cin, cout, trainable in self.conv_specs:
            if not trainable:
                continue
            if first_trainable:  # its input is the frozen part: no data-gradient needed
                first_trainable = False
                continue
            dst = self._dg(name, (cin, 9 * cout))
            call("ptb200_transpose_pack_f16", self.view(name + ".weight"), dst, cout, cin, 9, 1, 9 * cout)
        C = self.C
        dst = self._dg("rpn_conv", (C, 9 * C))
        call("ptb200_transpose_pack_f16", self.view("proposal_generator.rpn_head.conv.weight"), dst, C, C, 9, 1, 9 * C)
        dst = self._dg("rpn_heads", (C, 128))
        call("ptb200_transpose_pack_f16", self.view("proposal_generator.rpn_head._heads.weight"), dst, 128, C, 1, 0, 128)
        fin = self.pooled * self.pooled * C
        dst = self._dg("fc1", (fin, self.fc_dim))
        call("ptb200_transpose_pack_f16", self.view("roi_heads.box_head.fc1.weight"), dst, self.fc_dim, fin, 1, 0, self.fc_dim)
        dst = self._dg("fc2", (self.fc_dim, self.fc_dim))
        call("ptb200_transpose_pack_f16", self.view("roi_heads.box_head.fc2.weight"), dst, self.fc_dim, self.fc_dim, 1, 0, self.fc_dim)
        dst = self._dg("pred", (self.fc_dim, 128))
        call("ptb200_transpose_pack_f16", self.view("roi_heads.box_predictor._heads.weight"), dst, 128, self.fc_dim, 1, 0, 128)

    # ------------------------------------------------------------------ init (SURVEY 8d synthetic weights)
    def init_synthetic(self, seed=0):
        """Random init with the reference's initialisers: c2_msra_fill for VGG convs (vgg.py:61-63; the
        unconditional vgg16_caffe.pth load of vgg.py:127-152 has no file to read here), N(0,0.01) RPN
        head, c2_xavier_fill box head, N(0,0.01)/N(0,0.001) predictor (fast_rcnn.py:164-169)."""
        g = torch.Generator().manual_seed(seed)
        sd = OrderedDict()
        for name, cin, cout, _ in self.conv_specs:
            sd[name + ".weight"] = torch.randn(cout, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cout * 9))
            sd[name + ".bias"] = torch.zeros(cout)
        C, A, K = self.C, self.A, self.K
        p = "proposal_generator.rpn_head."
        sd[p + "conv.weight"] = torch.randn(C, C, 3, 3, generator=g) * 0.01
        sd[p + "conv.bias"] = torch.zeros(C)
        sd[p + "objectness_logits.weight"] = torch.randn(A, C, 1, 1, generator=g) * 0.01
        sd[p + "objectness_logits.bias"] = torch.zeros(A)
        sd[p + "anchor_deltas.weight"] = torch.randn(A * 8, C, 1, 1, generator=g) * 0.01
        sd[p + "anchor_deltas.bias"] = torch.zeros(A * 8)
        if "proposal_generator.anchor_generator.anchor_0" in self.segments:
            # keep what DifferentiableAnchorGenerator.__init__ wrote from the model's own cfg
            cur = self.view("proposal_generator.anchor_generator.anchor_0").detach().cpu().clone()
            if not bool(cur.any()):  # a bare arena (no generator constructed on it): the default config's anchors
                from .config import get_cfg
                cur = torch.tensor(get_cfg().MODEL.ANCHOR_GENERATOR.ANCHOR[0])
            sd["proposal_generator.anchor_generator.anchor_0"] = cur
        fin = C * self.pooled ** 2
        for name, fi, fo in (("roi_heads.box_head.fc1", fin, self.fc_dim), ("roi_heads.box_head.fc2", self.fc_dim, self.fc_dim)):
            bound = math.sqrt(3.0 / fi)
            sd[name + ".weight"] = (torch.rand(fo, fi, generator=g) * 2 - 1) * bound
            sd[name + ".bias"] = torch.zeros(fo)
        p = "roi_heads.box_predictor."
        sd[p + "cls_score.weight"] = torch.randn(K + 1, self.fc_dim, generator=g) * 0.01
        sd[p + "cls_score.bias"] = torch.zeros(K + 1)
        sd[p + "bbox_pred.weight"] = torch.randn(K * 8, self.fc_dim, generator=g) * 0.001
        sd[p + "bbox_pred.bias"] = torch.zeros(K * 8)
        self.load_state_dict(sd)
        return sd
