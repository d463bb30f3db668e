"""PTrainer on the B200 path: the per-step loop of `pt/engine/trainer.py` (run_step :263-392,
pseudo-labelling :179-257, EMA :431-449, resize :557-590, clip_gradient :592-603) driving
`GuassianGeneralizedRCNN` through its reference call signature. Host orchestration stays Python;
EMA, gradient clipping and the SGD step are single fused passes over the flat arenas, `resize`
runs on the device, and the data-parallel gradient all-reduce goes through torch.distributed
(NCCL over NVLink) on the flat gradient arena. Hooks / evaluation / checkpoint writers of the
reference trainer are out of scope."""
import random

import torch
import torch.distributed as dist

from .._lib import call, refresh_stream
from ..modeling.meta_arch.rcnn import build_model
from ..structures import Boxes, FreeInstances


def warmup_multistep_lr(base_lr, it, steps, gamma, warmup_factor, warmup_iters, warmup_method="linear"):
    """detectron2 WarmupMultiStepLR (configs/pt/final_c2f.yaml:6)."""
    f = 1.0
    if it < warmup_iters:
        if warmup_method == "constant":
            f = warmup_factor
        else:
            alpha = it / warmup_iters
            f = warmup_factor * (1 - alpha) + alpha
    k = sum(1 for s in steps if it >= s)
    return base_lr * f * (gamma ** k)


class PTrainer:
    def __init__(self, cfg, data_loader_iter, device=None, seed=0, loss_scale=1024.0):
        self.cfg = cfg
        self.device = torch.device(device or "cuda")
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.model = build_model(cfg, self.device, loss_scale=loss_scale)
        self.model_teacher = build_model(cfg, self.device, loss_scale=loss_scale, with_grads=False)
        self.model.init_synthetic(seed)
        if self.world > 1:  # DDP _sync_params_and_buffers (trainer.py:491-496)
            dist.broadcast(self.model.arena.data, src=0)
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

    # ------------------------------------------------------------------ EMA (trainer.py:431-449)
    @torch.no_grad()
    def _update_teacher_model(self, keep_rate=0.996):
        s, t = self.model.arena, self.model_teacher.arena
        call("ptb200_ema_update", t.data, s.data, t.total, float(keep_rate))
        t.pack()

    # ------------------------------------------------------------------ clip + SGD (trainer.py:383-386,592-603)
    def _optimizer_step(self, clip_norm=10.0):
        a = self.model.arena
        n = a.grads.numel()
        pre = 1.0 / self.world
        if self.world > 1:
            dist.all_reduce(a.grads)
        lr = warmup_multistep_lr(self.cfg.SOLVER.BASE_LR, self.iter, self.cfg.SOLVER.STEPS, self.cfg.SOLVER.GAMMA,
                                 self.cfg.SOLVER.WARMUP_FACTOR, self.cfg.SOLVER.WARMUP_ITERS,
                                 self.cfg.SOLVER.WARMUP_METHOD)
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
        self._optimizer_step(10.0)
        self.last_losses = {k: v.detach() for k, v in record_dict.items()}
        self.iter += 1
        return self.last_losses

    def train(self, num_iters):
        for _ in range(num_iters):
            self.run_step()
