"""PTrainer on the B200 path: the per-step loop of `pt/engine/trainer.py` (run_step :263-392,
pseudo-labelling :179-257, EMA :431-449, resize :557-590, clip_gradient :592-603) driving
`GuassianGeneralizedRCNN` through its reference call signature. Host orchestration stays Python;
EMA, gradient clipping and the SGD step are single fused passes over the flat arenas, `resize`
runs on the device, and the data-parallel gradient all-reduce goes through torch.distributed
(NCCL over NVLink) on the flat gradient arena. Hooks / evaluation / checkpoint writers of the
reference trainer are out of scope."""
import os
import random

import torch
import torch.distributed as dist

from . import dp
from .._lib import call, refresh_stream
from ..modeling.meta_arch.rcnn import build_model
from ..solver import lr_at_iter
from ..structures import Boxes, FreeInstances


class PTrainer:
    def __init__(self, cfg, data_loader_iter, device=None, seed=0, loss_scale=1024.0, use_cuda_graph=False,
                 graph_warmup=3, gt_capacity=64, concurrent=False, precision="f16"):
        """precision: "f16x3" = fp32-equivalent forward + backward (meets the 1e-3 parity with the reference's fp32
        step), "f16" = fp16 operands with fp32 accumulation / master weights (throughput mode)."""
        self.cfg = cfg
        self.precision = precision
        self.device = torch.device(device or "cuda")
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.model = build_model(cfg, self.device, loss_scale=loss_scale, precision=precision)
        self.model_teacher = build_model(cfg, self.device, loss_scale=loss_scale, with_grads=False,
                                         precision=precision)
        self.model.init_synthetic(seed)
        if self.world > 1:  # DDP _sync_params_and_buffers (trainer.py:491-496)
            dp.broadcast_params(self.model.arena.data, src=0)
            self.model.arena.pack()
        self.model_teacher.arena.data.copy_(self.model.arena.data)
        self.model_teacher.arena.pack()
        self.model.train()
        self.model_teacher.train()  # the teacher stays in train mode (trainer.py:302-313)
        self._data_loader_iter = data_loader_iter
        self.iter = 0
        self.start_iter = 0
        self.max_iter = cfg.SOLVER.MAX_ITER
        self.rng = random.Random(seed + 17 * self.rank)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._pix = [int(x) for x in cfg.MODEL.PIXEL_MEAN]
        self.last_losses = None
        # CUDA-graph mode: the whole forward/backward part of the step is captured once and replayed;
        # inputs are staged into persistent device buffers, only the all-reduce + optimizer stay eager
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._graph_warmup = graph_warmup
        self._gt_capacity = gt_capacity
        self._static = None
        # concurrent mode (graph step only): teacher pass, supervised pass and unsupervised pass are issued
        # on three streams (parallel branches of the captured graph), so the latency-bound proposal kernels
        # of one pass overlap the GEMMs of the others (leaving SMs free for them did not pay: 140 / 144 / 148
        # GEMM CTAs measured 14.55 / 14.54 / 14.42 ms per step)
        self.concurrent = concurrent
        self._streams = None
        self._comm_stream = None
        self._copy_stream = None
        self._prefetched = None
        self._slot = 0
        self._grads_reduced = False
        # world > 1 + concurrent graph step: the gradients behind the VGG backbone (29.3 M of 43.8 M floats) are
        # all-reduced INSIDE the graph on a communication stream while the backbone backward still runs, the backbone
        # bucket after the join (torch DDP's bucketed overlap, pt/engine/trainer.py:92-95). Round 1 recorded a "hang at
        # full size": every step ran, the process then blocked in dist.destroy_process_group() because the live graph
        # still held captured NCCL kernels -- call release_graphs() before tearing the process group down (bench.py and
        # tools/check_ddp.py do). Measured (profiles/r2b_allreduce_overlap_ab.json): 8 GPUs 37.18 -> 36.94 ms (f16x3),
        # 14.74 -> 14.69 ms (f16); nothing at 2 GPUs. PTB200_OVERLAP_ALLREDUCE=0 selects the single eager all-reduce.
        self.overlap_allreduce = os.environ.get("PTB200_OVERLAP_ALLREDUCE", "1") == "1"
        self.concurrent_gemm_ctas = int(os.environ.get("PTB200_GEMM_CTAS", "0"))  # 0 = one CTA per SM

    # ------------------------------------------------------------------ pseudo-labelling (trainer.py:179-257)
    def threshold_bbox(self, proposal_bbox_inst, proposal_type="roih"):
        image_shape = proposal_bbox_inst.image_size
        new = FreeInstances(image_shape)
        if proposal_type == "rpn":
            new.gt_boxes = Boxes(proposal_bbox_inst.proposal_boxes.tensor)
            new.objectness_logits = proposal_bbox_inst.objectness_logits
            new.pseudo_boxes = Boxes(proposal_bbox_inst.proposal_boxes.tensor)
        elif proposal_type == "roih":
            new.pseudo_boxes = Boxes(proposal_bbox_inst.pred_boxes.tensor)
            new.scores_logists = proposal_bbox_inst.scores_logists
            if proposal_bbox_inst.has("boxes_sigma"):
                new.boxes_sigma = proposal_bbox_inst.boxes_sigma
        new._count = proposal_bbox_inst._count
        return new

    def process_pseudo_label(self, proposals, proposal_type, psedo_label_method=""):
        if psedo_label_method != "all":
            raise ValueError("Unkown pseudo label boxes methods")
        return [self.threshold_bbox(p, proposal_type) for p in proposals], None

    @staticmethod
    def remove_label(label_data):
        for d in label_data:
            d.pop("instances", None)
        return label_data

    @staticmethod
    def add_label(unlabled_data, label):
        for d, inst in zip(unlabled_data, label):
            d["instances"] = inst
        return unlabled_data

    # ------------------------------------------------------------------ resize (trainer.py:557-590)
    def resize(self, data):
        out = []
        for d in data:
            img = d["image"]
            h, w = img.shape[-2], img.shape[-1]
            ratio = self.rng.uniform(0.5, 1.0)
            d_h, d_w = int(h * ratio), int(w * ratio)
            x1 = int((w - d_w) / 2)
            y1 = int((h - d_h) / 2)
            src = img if img.is_cuda else img.to(self.device, non_blocking=True)
            src = src.contiguous()
            dst = torch.empty_like(src)
            call("ptb200_resize_paste_u8", src, dst, h, w, d_h, d_w, x1, y1, self._pix[0], self._pix[1], self._pix[2])
            nd = dict(d)
            nd["image"] = dst
            inst = d["instances"]
            ni = FreeInstances(inst.image_size)
            ni._count = getattr(inst, "_count", None)
            for k, v in inst.get_fields().items():
                if k in ("gt_boxes", "pseudo_boxes"):
                    t = v.tensor * ratio
                    t[:, 0::2] += x1
                    t[:, 1::2] += y1
                    v = Boxes(t)
                ni.set(k, v)
            nd["instances"] = ni
            out.append(nd)
        return out

    def resize_dev(self, data, params_dev, ratio_dev):
        """`resize` with the random geometry held in device memory: params_dev int32 [n, 4] =
        (d_h, d_w, x1, y1), ratio_dev float32 [n]. Used by the CUDA-graph step (replay-safe)."""
        out = []
        for k, d in enumerate(data):
            img = d["image"]
            h, w = img.shape[-2], img.shape[-1]
            dst = torch.empty_like(img)
            call("ptb200_resize_paste_u8_dev", img, dst, h, w, params_dev[k], self._pix[0], self._pix[1], self._pix[2])
            nd = dict(d)
            nd["image"] = dst
            inst = d["instances"]
            ni = FreeInstances(inst.image_size)
            ni._count = getattr(inst, "_count", None)
            shift = params_dev[k, 2:4].to(torch.float32).repeat(2)  # (x1, y1, x1, y1)
            for key, v in inst.get_fields().items():
                if key in ("gt_boxes", "pseudo_boxes"):
                    v = Boxes(v.tensor * ratio_dev[k] + shift)
                ni.set(key, v)
            nd["instances"] = ni
            out.append(nd)
        return out

    # ------------------------------------------------------------------ EMA (trainer.py:431-449)
    @torch.no_grad()
    def _update_teacher_model(self, keep_rate=0.996, keep_dev=None):
        """keep_dev (device fp32 scalar) replaces keep_rate in the captured step: see `_stage`."""
        s, t = self.model.arena, self.model_teacher.arena
        if keep_dev is not None:
            call("ptb200_ema_update_dev", t.data, s.data, t.total, keep_dev)
        else:
            call("ptb200_ema_update", t.data, s.data, t.total, float(keep_rate))
        t.pack()

    # ------------------------------------------------------------------ clip + SGD (trainer.py:383-386,592-603)
    METRIC_KEYS = ("loss_cls_sup", "loss_box_reg_sup", "loss_rpn_cls_sup", "loss_rpn_loc_sup", "loss_cls_unsup",
                   "loss_box_reg_unsup", "loss_rpn_cls_unsup", "loss_rpn_loc_unsup")

    def _stash_metrics(self, losses):
        """Writes the step's loss scalars into the metrics tail of the gradient buffer (device copy, no sync) so that
        they are summed over ranks by the gradient all-reduce."""
        a = self.model.arena
        keys = [k for k in self.METRIC_KEYS if k in losses] or sorted(losses)
        self._metric_keys = keys
        a.metrics_tail.zero_()
        a.metrics_tail[:len(keys)].copy_(torch.stack([losses[k].detach().reshape(()).float() for k in keys]))

    def reduced_metrics(self):
        """The last step's losses averaged over the data-parallel ranks -- what the reference's `_write_metrics` logs
        on rank 0 (`comm.gather` + mean, pt/engine/trainer.py:394-429). Device tensors; reading them synchronises."""
        a = self.model.arena
        keys = getattr(self, "_metric_keys", None)
        if not keys:
            return {}
        vals = a.metrics_tail[:len(keys)] * dp.pre_scale()
        return {k: vals[i] for i, k in enumerate(keys)}

    def _optimizer_step(self, clip_norm=10.0, reduced=False):
        a = self.model.arena
        n = a.grads.numel()
        pre = dp.pre_scale()  # SUM all-reduce, 1/world folded into the clip / SGD kernels: DDP's gradient averaging
        if self.last_losses:
            self._stash_metrics(self.last_losses)
        if not reduced:
            # one collective over the whole arena + metrics tail (no-op at world 1)
            dp.allreduce_grads(a.grads_ext, bucket_elems=a.grads_ext.numel())
        elif self.world > 1:
            dist.all_reduce(a.metrics_tail)
        lr = lr_at_iter(self.cfg, self.iter)  # pt/solver/build.py: WarmupMultiStepLR unless the config says otherwise
        call("ptb200_grad_sumsq", a.grads, n, pre, self._sumsq)
        call("ptb200_clip_sgd_step", a.data[a.trainable_start:], a.grads, a.momentum, n, float(lr),
             float(self.cfg.SOLVER.MOMENTUM), float(self.cfg.SOLVER.WEIGHT_DECAY), float(clip_norm), pre,
             self._sumsq)
        a.pack()

    # ------------------------------------------------------------------ the step (trainer.py:263-392)
    def run_step(self):
        assert self.model.training, "[PTrainer] model was changed to eval mode!"
        refresh_stream()
        cfg = self.cfg
        label_data_q, label_data_k, unlabel_data_q, unlabel_data_k = next(self._data_loader_iter)
        record_dict = {}
        self.model.zero_grad()
        if self.iter < cfg.UNSUPNET.BURN_UP_STEP:
            label_data_q = self.resize(list(label_data_q) + list(label_data_k))
            rec, _, _, _ = self.model(label_data_q, branch="supervised")
            record_dict.update(rec)
            loss_dict = {k: v * 1.0 for k, v in rec.items() if k[:4] == "loss"}
        else:
            if self.iter == cfg.UNSUPNET.BURN_UP_STEP:
                self._update_teacher_model(keep_rate=0.00)
            elif (self.iter - cfg.UNSUPNET.BURN_UP_STEP) % cfg.UNSUPNET.TEACHER_UPDATE_ITER == 0:
                self._update_teacher_model(keep_rate=cfg.UNSUPNET.EMA_KEEP_RATE)
            with torch.no_grad():
                _, _, proposals_roih_unsup_k, _ = self.model_teacher(unlabel_data_k, branch="unsup_data_weak")
            pseudo, _ = self.process_pseudo_label(proposals_roih_unsup_k, "roih", "all")
            unlabel_data_q = self.add_label(self.remove_label([dict(d) for d in unlabel_data_q]), pseudo)
            unlabel_data_q = self.resize(unlabel_data_q)
            label_data_q = self.resize(label_data_q)
            all_label_data = label_data_q + list(label_data_k)
            rec_l, _, _, _ = self.model(all_label_data, branch="supervised")
            for k, v in rec_l.items():
                record_dict[k + "_sup"] = v
            rec_u, _, _, _ = self.model(unlabel_data_q, branch="unsupervised", danchor=True)
            for k, v in rec_u.items():
                record_dict[k + "_unsup"] = v
            loss_dict = {}
            for k, v in record_dict.items():
                if k[:4] == "loss":
                    if k.split("_")[-1] == "sup":
                        loss_dict[k] = v * cfg.UNSUPNET.SOURCE_LOSS_WEIGHT
                    elif k.split("_")[-1] == "unsup":
                        loss_dict[k] = v * cfg.UNSUPNET.TARGET_UNSUP_LOSS_WEIGHT
                    else:
                        raise NotImplementedError
        losses = sum(loss_dict.values())
        losses.backward()
        self.last_losses = {k: v.detach() for k, v in record_dict.items()}
        self._optimizer_step(10.0)
        self.iter += 1
        return self.last_losses

    # ------------------------------------------------------------------ CUDA-graph step
    def _make_static(self, data):
        """Persistent device buffers for one step's inputs (same shapes every step)."""
        lq, lk, uq, uk = data
        dev = self.device
        cap = self._gt_capacity
        st = {"groups": {}, "n": {}}
        for name, grp, labelled in (("lq", lq, True), ("lk", lk, True), ("uq", uq, False), ("uk", uk, False)):
            n = len(grp)
            h, w = grp[0]["image"].shape[-2:]
            g = {"images": torch.empty(n, 3, h, w, dtype=torch.uint8, device=dev),
                 # double-buffered landing zone of the host->device image copies (see _prefetch_images)
                 "staging": [torch.empty(n, 3, h, w, dtype=torch.uint8, device=dev) for _ in range(2)]}
            if labelled:
                g["gt_boxes"] = torch.zeros(n, cap, 4, dtype=torch.float32, device=dev)
                g["gt_boxes"][..., 2:] = 1.0
                g["gt_classes"] = torch.zeros(n, cap, dtype=torch.int32, device=dev)
                g["gt_count"] = torch.zeros(n, dtype=torch.int32, device=dev)
                # pinned host staging, one set per slot: the host runs ahead of the device (a graph replay is launched
                # in microseconds), so a buffer is rewritten only after the copies that read it have completed
                g["pin_boxes"] = [torch.zeros(n, cap, 4, dtype=torch.float32).pin_memory() for _ in range(2)]
                g["pin_classes"] = [torch.zeros(n, cap, dtype=torch.int32).pin_memory() for _ in range(2)]
                g["pin_count"] = [torch.zeros(n, dtype=torch.int32).pin_memory() for _ in range(2)]
            st["groups"][name] = g
        nq = len(lq) + len(uq)
        st["resize_params"] = torch.zeros(nq, 4, dtype=torch.int32, device=dev)
        st["resize_ratio"] = torch.ones(nq, dtype=torch.float32, device=dev)
        st["pin_params"] = [torch.zeros(nq, 4, dtype=torch.int32).pin_memory() for _ in range(2)]
        st["pin_ratio"] = [torch.ones(nq, dtype=torch.float32).pin_memory() for _ in range(2)]
        st["ema_keep"] = torch.ones(1, dtype=torch.float32, device=dev)
        st["pin_keep"] = [torch.ones(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        st["pin_done"] = [None, None]      # main-stream events: the host->device copies out of a slot's pinned buffers ran
        st["stage_ready"] = [None, None]   # copy-stream events: images of a slot have landed
        st["stage_free"] = [None, None]    # main-stream events: a slot's images were moved into the static buffers
        return st

    def _prefetch_images(self, data, slot):
        """Issues the host->device copies of one batch's images on a copy stream into staging slot `slot`.
        Called right after the graph of the CURRENT step has been launched, for the NEXT step's batch, so that the
        PCIe transfer (25.6 MB at 3x800x1333, ~1 ms) overlaps the current step's kernels."""
        st = self._static
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        cs = self._copy_stream
        if st["stage_free"][slot] is not None:
            cs.wait_event(st["stage_free"][slot])
        else:
            cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cs):
            for name, grp in zip(("lq", "lk", "uq", "uk"), data):
                buf = st["groups"][name]["staging"][slot]
                for k, d in enumerate(grp):
                    buf[k].copy_(d["image"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        st["stage_ready"][slot] = ev

    def _stage(self, data, slot):
        """Moves one step's inputs into the persistent device buffers the graph reads (outside the graph): the
        images from their staging slot (device-to-device, after the copy stream's event), ground truth and the
        resize geometry from pinned host buffers."""
        st = self._static
        cap = self._gt_capacity
        main = torch.cuda.current_stream()
        main.wait_event(st["stage_ready"][slot])
        if st["pin_done"][slot] is not None:
            # the copies issued two steps ago from this slot's pinned buffers must have run before the host
            # rewrites them (they have, unless the host is more than a full step ahead of the device)
            st["pin_done"][slot].synchronize()
        for name, grp in zip(("lq", "lk", "uq", "uk"), data):
            g = st["groups"][name]
            g["images"].copy_(g["staging"][slot], non_blocking=True)
            if "gt_boxes" in g:
                pb, pc, pn = g["pin_boxes"][slot], g["pin_classes"][slot], g["pin_count"][slot]
                pb.zero_()
                pb[..., 2:] = 1.0
                pc.zero_()
                for k, d in enumerate(grp):
                    inst = d["instances"]
                    m = len(inst.gt_boxes)
                    if m > cap:
                        raise ValueError(f"{m} gt boxes exceed gt_capacity={cap}")
                    pb[k, :m] = inst.gt_boxes.tensor
                    pc[k, :m] = inst.gt_classes.to(torch.int32)
                    pn[k] = m
                g["gt_boxes"].copy_(pb, non_blocking=True)
                g["gt_classes"].copy_(pc, non_blocking=True)
                g["gt_count"].copy_(pn, non_blocking=True)
        # PTrainer.resize geometry (trainer.py:561-566). Draw order as in run_step (unlabel_q first, then
        # label_q: trainer.py:333-334); storage order: label_q rows first, then unlabel_q rows.
        nl = len(data[0])
        for grp, base in ((data[2], nl), (data[0], 0)):
            for j, d in enumerate(grp):
                h, w = d["image"].shape[-2:]
                ratio = self.rng.uniform(0.5, 1.0)
                d_h, d_w = int(h * ratio), int(w * ratio)
                k = base + j
                pp, pr = st["pin_params"][slot], st["pin_ratio"][slot]
                pp[k, 0] = d_h
                pp[k, 1] = d_w
                pp[k, 2] = int((w - d_w) / 2)
                pp[k, 3] = int((h - d_h) / 2)
                pr[k] = ratio
        st["resize_params"].copy_(st["pin_params"][slot], non_blocking=True)
        st["resize_ratio"].copy_(st["pin_ratio"][slot], non_blocking=True)
        # teacher update cadence (trainer.py:296-298): the captured EMA reads its keep rate from device memory
        cfg = self.cfg
        update = (self.iter - cfg.UNSUPNET.BURN_UP_STEP) % cfg.UNSUPNET.TEACHER_UPDATE_ITER == 0
        st["pin_keep"][slot][0] = float(cfg.UNSUPNET.EMA_KEEP_RATE) if update else 1.0
        st["ema_keep"].copy_(st["pin_keep"][slot], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(main)
        st["stage_free"][slot] = ev
        st["pin_done"][slot] = ev

    def _static_batches(self):
        st = self._static
        out = {}
        for name, g in st["groups"].items():
            n, _, h, w = g["images"].shape
            batch = []
            for k in range(n):
                d = {"image": g["images"][k], "height": h, "width": w}
                if "gt_boxes" in g:
                    inst = FreeInstances((h, w), gt_boxes=Boxes(g["gt_boxes"][k]), gt_classes=g["gt_classes"][k])
                    inst._count = g["gt_count"][k]
                    d["instances"] = inst
                batch.append(d)
            out[name] = batch
        return out

    def _graph_body(self):
        """The capturable part of the post-burn-in step (pt/engine/trainer.py:291-384): EMA, teacher
        pass, pseudo labels, resize, both student passes, backward. No host<->device traffic."""
        cfg = self.cfg
        st = self._static
        b = self._static_batches()
        nl = len(b["lq"])
        # the C-ABI launches go to a cached stream handle: re-read it INSIDE the capture (torch captures on a side
        # stream). Without this the EMA + teacher re-pack below were issued eagerly on the outer stream while the graph
        # was being captured and never replayed (round 1; caught by test_graph_step_honours_teacher_update_iter)
        refresh_stream()
        self.model.zero_grad()
        self._update_teacher_model(keep_dev=st["ema_keep"])
        with torch.no_grad():
            _, _, roih, _ = self.model_teacher(b["uk"], branch="unsup_data_weak")
        pseudo, _ = self.process_pseudo_label(roih, "roih", "all")
        uq = self.add_label([dict(d) for d in b["uq"]], pseudo)
        uq = self.resize_dev(uq, st["resize_params"][nl:], st["resize_ratio"][nl:])
        lq = self.resize_dev(b["lq"], st["resize_params"][:nl], st["resize_ratio"][:nl])
        rec = {}
        rec_l, _, _, _ = self.model(lq + b["lk"], branch="supervised")
        for k, v in rec_l.items():
            rec[k + "_sup"] = v
        rec_u, _, _, _ = self.model(uq, branch="unsupervised", danchor=True)
        for k, v in rec_u.items():
            rec[k + "_unsup"] = v
        total = 0
        for k, v in rec.items():
            wgt = cfg.UNSUPNET.SOURCE_LOSS_WEIGHT if k.endswith("_sup") else cfg.UNSUPNET.TARGET_UNSUP_LOSS_WEIGHT
            total = total + v * wgt
        total.backward()
        return {k: v.detach() for k, v in rec.items()}

    def _resize_images_dev(self, data, params_dev):
        out = []
        for k, d in enumerate(data):
            img = d["image"]
            h, w = img.shape[-2], img.shape[-1]
            dst = torch.empty_like(img)
            call("ptb200_resize_paste_u8_dev", img, dst, h, w, params_dev[k], self._pix[0], self._pix[1], self._pix[2])
            nd = dict(d)
            nd["image"] = dst
            out.append(nd)
        return out

    def _resize_instances_dev(self, instances, params_dev, ratio_dev):
        out = []
        for k, inst in enumerate(instances):
            ni = FreeInstances(inst.image_size)
            ni._count = getattr(inst, "_count", None)
            shift = params_dev[k, 2:4].to(torch.float32).repeat(2)
            for key, v in inst.get_fields().items():
                if key in ("gt_boxes", "pseudo_boxes"):
                    v = Boxes(v.tensor * ratio_dev[k] + shift)
                ni.set(key, v)
            out.append(ni)
        return out

    def _graph_body_concurrent(self):
        """Same computation as `_graph_body`, issued as three parallel branches (teacher forward, supervised
        forward + backward, unsupervised forward + backward) that join before the optimizer step."""
        from .. import ops
        cfg = self.cfg
        st = self._static
        b = self._static_batches()
        nl = len(b["lq"])
        main = torch.cuda.current_stream()
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(3)]
        s_t, s_1, s_2 = self._streams
        ops.GEMM_MAX_CTAS[0] = self.concurrent_gemm_ctas
        self.model.zero_grad()
        for s in (s_t, s_1, s_2):
            s.wait_stream(main)
        overlap = self.world > 1 and self.overlap_allreduce
        head_events = []
        if overlap:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=self.device)

            def hook():
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
                head_events.append(ev)
            self.model.heads_backward_hook = hook
        e_pseudo = torch.cuda.Event()
        keep = []  # tensors that cross streams stay referenced until the join
        with torch.cuda.stream(s_t):
            refresh_stream()
            self._update_teacher_model(keep_dev=st["ema_keep"])
            with torch.no_grad():
                _, _, roih, _ = self.model_teacher(b["uk"], branch="unsup_data_weak")
            pseudo, _ = self.process_pseudo_label(roih, "roih", "all")
            e_pseudo.record(s_t)
            keep.append((roih, pseudo))
        rec = {}
        w_sup, w_unsup = cfg.UNSUPNET.SOURCE_LOSS_WEIGHT, cfg.UNSUPNET.TARGET_UNSUP_LOSS_WEIGHT
        # Each student pass runs its backward on its own branch as soon as its forward is done (the weighted
        # sum of trainer.py:364-381 is linear, and both backward chains accumulate into the gradient arena
        # with fp32 atomics): the supervised backward overlaps the latency-bound proposal / NMS kernels of the
        # teacher and unsupervised passes instead of waiting for the join.
        with torch.cuda.stream(s_1):
            refresh_stream()
            lq = self.resize_dev(b["lq"], st["resize_params"][:nl], st["resize_ratio"][:nl])
            rec_l, _, _, _ = self.model(lq + b["lk"], branch="supervised")
            keep.append(lq)
            sum(v * w_sup for v in rec_l.values()).backward()
            refresh_stream()
        with torch.cuda.stream(s_2):
            refresh_stream()
            uq_img = self._resize_images_dev(b["uq"], st["resize_params"][nl:])

            def provider():
                torch.cuda.current_stream().wait_event(e_pseudo)
                return self._resize_instances_dev(pseudo, st["resize_params"][nl:], st["resize_ratio"][nl:])
            rec_u, _, _, _ = self.model(uq_img, branch="unsupervised", danchor=True, targets_provider=provider)
            keep.append(uq_img)
            sum(v * w_unsup for v in rec_u.values()).backward()
            refresh_stream()
        if overlap:
            # Gradient all-reduce inside the captured step, in two segments of the flat arena: everything behind
            # the VGG backbone (29.3 M of 43.8 M floats: fc1 alone is 103 MB) is reduced on a communication
            # stream as soon as BOTH student passes have finished their head backward, overlapping the backbone
            # data / weight gradient GEMMs; the backbone segment follows after the join. (torch DDP's bucketed
            # overlap at pt/engine/trainer.py:92-95, with two buckets cut at the arena's natural boundary.)
            self.model.heads_backward_hook = None
            g = self.model.arena.grads
            cut = self.model.arena.head_bucket_start()
            assert len(head_events) == 2
            sc = self._comm_stream
            sc.wait_stream(main)
            for ev in head_events:
                sc.wait_event(ev)
            with torch.cuda.stream(sc):
                dist.all_reduce(g[cut:])
        for s in (s_t, s_1, s_2):
            main.wait_stream(s)
        refresh_stream()
        if overlap:
            # the second collective must be ORDERED after the first one: synchronous collectives are issued on
            # the caller's stream, and two collectives of one communicator running concurrently from two streams
            # (in an order that may differ between ranks) dead-lock NCCL
            main.wait_stream(sc)
            dist.all_reduce(g[:cut])
            self._grads_reduced = True
        for k, v in rec_l.items():
            rec[k + "_sup"] = v
        for k, v in rec_u.items():
            rec[k + "_unsup"] = v
        ops.GEMM_MAX_CTAS[0] = 0
        self._keep = keep
        return {k: v.detach() for k, v in rec.items()}

    def run_step_graphed(self):
        """Post-burn-in step with the forward/backward part replayed from a CUDA graph."""
        assert self.iter > self.cfg.UNSUPNET.BURN_UP_STEP, "graph mode covers the steady-state (EMA) iterations"
        refresh_stream()
        if self._prefetched is None:
            data = next(self._data_loader_iter)
            if self._static is None:
                self._static = self._make_static(data)
            self._slot = 0
            self._prefetch_images(data, self._slot)
        else:
            data = self._prefetched
        slot = self._slot
        self._stage(data, slot)
        body = self._graph_body_concurrent if self.concurrent else self._graph_body
        if self._graph is None:
            if self._graph_warmup > 0:  # eager warm-up on the static buffers (allocator, lazy inits)
                self._graph_warmup -= 1
                self.last_losses = body()
            else:
                torch.cuda.synchronize()
                self._graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self._graph):
                    self._graph_losses = body()
                refresh_stream()
                self._graph.replay()
                self.last_losses = self._graph_losses
        else:
            self._graph.replay()
            self.last_losses = self._graph_losses
        refresh_stream()
        # the next batch's images travel host->device while this step's kernels run
        self._prefetched = next(self._data_loader_iter)
        self._slot = 1 - slot
        self._prefetch_images(self._prefetched, self._slot)
        self._optimizer_step(10.0, reduced=self._grads_reduced)
        self.iter += 1
        return self.last_losses

    # ------------------------------------------------------------------ checkpoints (trainer.py:104-111,466-495)
    def build_checkpointer(self, save_dir=None):
        """`DetectionTSCheckpointer(EnsembleTSModel(teacher, student), cfg.OUTPUT_DIR, optimizer=...)`
        (trainer.py:104-111); rank 0 writes (fvcore `save_to_disk=comm.is_main_process()`)."""
        from ..checkpoint import ArenaSGDState, DetectionTSCheckpointer, EnsembleTSModel, SchedulerState
        ens = EnsembleTSModel(self.model_teacher, self.model)
        self.checkpointer = DetectionTSCheckpointer(ens, self.cfg.OUTPUT_DIR if save_dir is None else save_dir,
                                                    save_to_disk=self.rank == 0, optimizer=ArenaSGDState(self),
                                                    scheduler=SchedulerState(self))
        return self.checkpointer

    def save_checkpoint(self, name=None):
        """What `hooks.PeriodicCheckpointer` writes every SOLVER.CHECKPOINT_PERIOD iterations (trainer.py:523-527):
        `model_{iter:07d}.pth` holding both detectors, the momentum buffers and the last finished iteration."""
        if getattr(self, "checkpointer", None) is None:
            self.build_checkpointer()
        last = self.iter - 1
        return self.checkpointer.save(name or "model_{:07d}".format(max(last, 0)), iteration=last)

    def resume_or_load(self, resume=False):
        """trainer.py:466-495. resume=True loads model + optimizer + scheduler from `cfg.MODEL.WEIGHTS` (the reference
        passes that path, not `last_checkpoint`: trainer.py:478-481) and continues at the stored iteration + 1; when
        MODEL.WEIGHTS is empty, `<OUTPUT_DIR>/last_checkpoint` is used (what the reference's docstring describes).
        resume=False loads only the weights of cfg.MODEL.WEIGHTS and starts from iteration 0. Under data parallelism
        rank 0's arenas are broadcast afterwards (DDP `_sync_params_and_buffers`, :491-494)."""
        if getattr(self, "checkpointer", None) is None:
            self.build_checkpointer()
        if resume:
            path = self.cfg.MODEL.WEIGHTS
            if not path and self.checkpointer.has_checkpoint():
                path = self.checkpointer.get_checkpoint_file()
            rest = self.checkpointer.load(path, checkpointables=["optimizer", "scheduler"])
            self.start_iter = rest.get("iteration", -1) + 1
            self.iter = self.start_iter
        else:
            self.checkpointer.load(self.cfg.MODEL.WEIGHTS, checkpointables=[])
        if self.world > 1:
            it = torch.tensor([self.iter], dtype=torch.int64, device=self.device)
            dist.broadcast(it, src=0)
            self.iter = self.start_iter = int(it.item())
            for a in (self.model.arena, self.model_teacher.arena):
                dist.broadcast(a.data, src=0)
                a.pack()
            dist.broadcast(self.model.arena.momentum, src=0)
        # the captured graph holds no parameter values (it reads the arenas), so it stays valid
        return self.start_iter

    def release_graphs(self):
        """Drops the captured step graph (and what it keeps alive). Call before `dist.destroy_process_group()` when the
        graph holds captured NCCL kernels (in-graph gradient all-reduce): tearing the communicator down under a live
        graph blocks inside destroy_process_group (measured: every step of tools/check_ddp.py passed at 800x1333 and
        the process then hung there -- the "hang at full size" of round 1)."""
        import gc
        torch.cuda.synchronize(self.device)
        self._graph = None
        self._graph_losses = None
        self._keep = None
        self._prefetched = None
        gc.collect()
        torch.cuda.synchronize(self.device)

    def step(self):
        if self.use_cuda_graph and self.iter > self.cfg.UNSUPNET.BURN_UP_STEP:
            return self.run_step_graphed()
        return self.run_step()

    def check_finite(self):
        """Lazy divergence check (one host sync; NOT part of the step): the non-finite-proposal flags of both detectors
        (`proposal_utils.py:117-122` raises per image, on the host, in training) and the last step's losses. Raises
        FloatingPointError."""
        self.model.proposal_generator.raise_if_nonfinite()
        self.model_teacher.proposal_generator.raise_if_nonfinite()
        if self.last_losses:
            total = torch.stack([v.reshape(()).float() for v in self.last_losses.values()]).sum()
            if not bool(torch.isfinite(total)):
                raise FloatingPointError(f"Loss became infinite or NaN at iteration={self.iter - 1}!\n"
                                         f"loss_dict = { {k: float(v) for k, v in self.last_losses.items()} }")

    def train(self, num_iters=None, check_period=20):
        """`train_loop(start_iter, max_iter)` (trainer.py:154-176) without the d2 hook machinery: runs to
        cfg.SOLVER.MAX_ITER (or for `num_iters` iterations), checking for divergence every `check_period` iterations
        (the reference does it every iteration at the price of a host sync per image); when a checkpointer has been built
        (`build_checkpointer` / `resume_or_load`) a checkpoint is written every SOLVER.CHECKPOINT_PERIOD iterations
        and `model_final` at MAX_ITER, as `hooks.PeriodicCheckpointer` does (trainer.py:523-527)."""
        end = self.max_iter if num_iters is None else min(self.iter + num_iters, self.max_iter)
        period = int(self.cfg.SOLVER.CHECKPOINT_PERIOD)
        ck = getattr(self, "checkpointer", None)
        while self.iter < end:
            self.step()
            if check_period > 0 and (self.iter % check_period == 0 or self.iter >= end):
                self.check_finite()
            if ck is not None:
                if period > 0 and self.iter % period == 0:  # fvcore PeriodicCheckpointer.step: (iteration + 1) % period
                    self.save_checkpoint()
                if self.iter >= self.max_iter:
                    self.save_checkpoint("model_final")
        return self.last_losses
