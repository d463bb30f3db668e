"""CPU: the HOST ORCHESTRATION of `PTrainer.run_step` (the product code: burn-in branch, teacher copy / EMA cadence,
pseudo-label repack, the order of the `resize` draws, loss weighting, backward, optimizer step) executed against the
fixtures of the reference's own trainer -- with every device piece replaced by the oracle: the two detectors are
oracle models behind the package's input / output containers, `resize` / EMA / clip + SGD are the oracle's functions.
What runs unmodified is `PTrainer.run_step`, `process_pseudo_label`, `threshold_bbox`, `add_label`, `remove_label` and
`solver.lr_at_iter`; since all arithmetic is the oracle's, losses and parameters must match the reference trainer to
the oracle's own tolerance. (The same steps on the CUDA kernels: tests/test_trainer_step_gpu.py,
tests/test_zz_next_rows_gpu.py.)"""
import os
import random

import pytest
import torch

from oracle import pt_oracle as O
from probabilisticteacher_b200.engine import trainer as trainer_mod
from probabilisticteacher_b200.config import c2f_config
from probabilisticteacher_b200.engine.trainer import PTrainer
from probabilisticteacher_b200.solver import lr_at_iter
from probabilisticteacher_b200.structures import Boxes, FreeInstances

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sample_idx(numel, n=64):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


class _Sampler:
    def __init__(self, pr):
        self.pr = pr

    def prio(self, tag, n):
        grp, which = tag[0].split("_")
        return self.pr[grp][0 if which == "pos" else 1][tag[1]][:n]


class _Detector:
    """An oracle model behind the boundary of GuassianGeneralizedRCNN: package containers in, package containers out."""

    def __init__(self, om):
        self.om = om
        self.training = True

    @staticmethod
    def _to_oracle(batch):
        out = []
        for d in batch:
            nd = {"image": d["image"], "height": d.get("height"), "width": d.get("width")}
            if "instances" in d:
                i = d["instances"]
                n = None if getattr(i, "_count", None) is None else int(i._count)
                f = {}
                for k, v in i.get_fields().items():
                    v = v.tensor if hasattr(v, "tensor") else v
                    v = v if n is None else v[:n]
                    f[k] = O.OBoxes(v) if k.endswith("boxes") else v
                nd["instances"] = O.OInst(tuple(i.image_size), **f)
            out.append(nd)
        return out

    def __call__(self, batch, branch="supervised", danchor=False):
        losses, _, roih, _ = self.om(self._to_oracle(batch), branch=branch, danchor=danchor)
        if branch != "unsup_data_weak":
            return losses, [], [], None
        out = []
        for r in roih:
            inst = FreeInstances(tuple(r.image_size), pred_boxes=Boxes(O._bt(r.pred_boxes)), scores=r.scores,
                                 pred_classes=r.pred_classes, scores_logists=r.scores_logists, boxes_sigma=r.boxes_sigma)
            inst._count = torch.tensor(len(r.scores))
            out.append(inst)
        return {}, [], out, None

    def zero_grad(self):
        self.om.zero_grad()


class _HostTrainer(PTrainer):
    """PTrainer with the device pieces swapped for the oracle's; `run_step` and the pseudo-label methods are inherited."""

    def __init__(self, cfg, ocfg, loader, student, teacher):  # no CUDA models are built
        self.cfg, self.ocfg = cfg, ocfg
        self.model, self.model_teacher = _Detector(student), _Detector(teacher)
        self._data_loader_iter = loader
        self.iter, self.world, self.rank = 0, 1, 0
        self.rng = random.Random(0)
        self.last_losses = None
        self.opt = O.make_optimizer(student, ocfg)

    def resize(self, data):
        ratios = [self.rng.uniform(0.5, 1.0) for _ in data]  # one draw per image, in order (trainer.py:561)
        out = O.resize_batch(_Detector._to_oracle(data), ratios, self.model.om.pixel_mean.flatten())
        res = []
        for d in out:
            i = d["instances"]
            f = {k: (Boxes(O._bt(getattr(i, k))) if k.endswith("boxes") else getattr(i, k))
                 for k in ("gt_boxes", "gt_classes", "pseudo_boxes", "scores_logists", "boxes_sigma") if i.has(k)}
            res.append({"image": d["image"], "height": d["height"], "width": d["width"],
                        "instances": FreeInstances(tuple(i.image_size), **f)})
        return res

    def _update_teacher_model(self, keep_rate=0.996):
        O.ema_update(self.model_teacher.om, self.model.om, keep_rate)

    def _optimizer_step(self, clip_norm=10.0, reduced=False):
        for grp in self.opt.param_groups:
            grp["lr"] = lr_at_iter(self.cfg, self.iter)
        O.clip_gradient(self.model.om.parameters(), clip_norm)
        self.opt.step()


@pytest.fixture(autouse=True)
def _no_device(monkeypatch):
    """`run_step` starts by re-reading torch's current CUDA stream for the kernel launches: there are none here."""
    monkeypatch.setattr(trainer_mod, "refresh_stream", lambda: None)


class _Ratios:
    def __init__(self, draws):
        self.draws = list(draws)

    def uniform(self, a, b):
        return self.draws.pop(0)


def _check_params(model, ref, what):
    sd = model.ref_state_dict()
    for k, v in ref.items():
        mine = sd[k].detach().reshape(-1)[_sample_idx(sd[k].numel())]
        err = float((mine - v).abs().max())
        assert err <= 2e-6 + 2e-5 * float(v.abs().max()), (what, k, err)


def _run(fixture, cfg, ocfg, batches, n_teacher_checks=True):
    G = torch.load(os.path.join(GOLD, fixture), weights_only=False)
    student = O.OracleRCNN(ocfg, seed=G["seed"])
    teacher = O.OracleRCNN(ocfg, seed=G["teacher_seed"])
    student.sampler = _Sampler(G["prio"])
    tr = _HostTrainer(cfg, ocfg, batches(G), student, teacher)
    for it, ref in enumerate(G["steps"]):
        tr.rng = _Ratios(ref["ratios"])
        losses = tr.run_step()
        assert tr.rng.draws == [] and tr.iter == it + 1
        got = {k: float(v) for k, v in losses.items()}
        assert set(got) == set(ref["losses"])
        for k, v in ref["losses"].items():
            assert abs(got[k] - v) <= 2e-5 * max(abs(v), 1e-6), (it, k, got[k], v)
        _check_params(student, ref["student"], ("student", it))
        if "teacher" in ref:
            _check_params(teacher, ref["teacher"], ("teacher", it))
    return G


def _post_burn_in_batches(G):
    H, W = G["H"], G["W"]

    def batch():
        lab = [{"image": im.clone(), "height": H, "width": W,
                "instances": FreeInstances((H, W), gt_boxes=Boxes(b.clone()), gt_classes=c.clone())}
               for im, b, c in zip(G["lab_images"], G["gt_boxes"], G["gt_classes"])]
        unl = [{"image": im.clone(), "height": H, "width": W} for im in G["unl_images"]]
        return lab, unl
    while True:
        lab, unl = batch()
        lab_k, _ = batch()
        _, unl_k = batch()
        yield lab, lab_k, unl, unl_k


def test_run_step_host_logic_two_post_burn_in_steps():
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.SOLVER.WARMUP_ITERS = 0
    _run("pt_reference_step_golden.pt", cfg, O.OracleCfg(num_classes=8), _post_burn_in_batches)


def test_run_step_host_logic_trainer_hyper_parameters_off_their_defaults():
    G = torch.load(os.path.join(GOLD, "pt_reference_step_oddcfg_golden.pt"), weights_only=False)
    t = G["trainer_cfg"]
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.UNSUPNET.SOURCE_LOSS_WEIGHT, cfg.UNSUPNET.TARGET_UNSUP_LOSS_WEIGHT = t["source_loss_weight"], t["target_unsup_loss_weight"]
    cfg.UNSUPNET.EMA_KEEP_RATE, cfg.UNSUPNET.TEACHER_UPDATE_ITER = t["ema_keep_rate"], t["teacher_update_iter"]
    cfg.SOLVER.MOMENTUM, cfg.SOLVER.WEIGHT_DECAY, cfg.SOLVER.BASE_LR, cfg.SOLVER.WARMUP_ITERS = \
        t["momentum"], t["weight_decay"], t["base_lr"], 0
    ocfg = O.OracleCfg(num_classes=8, base_lr=t["base_lr"], momentum=t["momentum"], weight_decay=t["weight_decay"])
    _run("pt_reference_step_oddcfg_golden.pt", cfg, ocfg, _post_burn_in_batches)


def test_run_step_host_logic_burn_in():
    def batches(G):
        H, W = G["H"], G["W"]

        def view(tag):
            return [{"image": im.clone(), "height": H, "width": W,
                     "instances": FreeInstances((H, W), gt_boxes=Boxes(b.clone()), gt_classes=c.clone())}
                    for im, b, c in zip(G[f"lab_{tag}_images"], G[f"gt_boxes_{tag}"], G[f"gt_classes_{tag}"])]
        while True:
            yield view("q"), view("k"), [], []
    cfg = c2f_config()
    cfg.SOLVER.WARMUP_ITERS = 0
    assert cfg.UNSUPNET.BURN_UP_STEP > 2
    _run("pt_reference_burnin_golden.pt", cfg, O.OracleCfg(num_classes=8), batches)


def test_resize_host_bookkeeping(monkeypatch):
    """`PTrainer.resize` / `resize_dev` with the kernel behind them emulated in torch (F.interpolate + truncation, what
    the oracle's restatement of pt/engine/trainer.py:557-590 does): the geometry handed to the kernel, the box
    scaling / shifting, the untouched inputs and the carried fields are the product's host code."""
    import torch.nn.functional as F

    def fake_call(name, src, dst, h, w, *rest):
        if name == "ptb200_resize_paste_u8_dev":
            params, m = rest[0], rest[1:4]
            d_h, d_w, x1, y1 = (int(v) for v in params)
        else:
            assert name == "ptb200_resize_paste_u8"
            d_h, d_w, x1, y1 = rest[:4]
            m = rest[4:7]
        dst.copy_(torch.tensor(m, dtype=torch.uint8).view(3, 1, 1).expand_as(dst))
        dst[:, y1:y1 + d_h, x1:x1 + d_w] = F.interpolate(src.unsqueeze(0).float(), size=(d_h, d_w), align_corners=False,
                                                         mode="bilinear").squeeze(0).to(torch.uint8)
    monkeypatch.setattr(trainer_mod, "call", fake_call)
    mean = torch.tensor([103.53, 116.28, 123.675])
    tr = PTrainer.__new__(PTrainer)
    tr.device, tr._pix = torch.device("cpu"), [int(x) for x in mean]
    H, W, ratios = 96, 131, [0.5, 0.8125, 1.0]
    batch = O.synthetic_batch(3, H, W, 8, 5)
    ref = O.resize_batch(batch, ratios, mean)

    def mine():
        return [{"image": d["image"].clone(), "height": H, "width": W, "file_name": f"img{k}",
                 "instances": FreeInstances((H, W), gt_boxes=Boxes(d["instances"].gt_boxes.tensor.clone()),
                                            gt_classes=d["instances"].gt_classes.clone())} for k, d in enumerate(batch)]
    tr.rng = _Ratios(ratios)
    data = mine()
    host = tr.resize(data)
    params = torch.tensor([[int(H * r), int(W * r), int((W - int(W * r)) / 2), int((H - int(H * r)) / 2)] for r in ratios],
                          dtype=torch.int32)
    dev = tr.resize_dev(mine(), params, torch.tensor(ratios))
    for k in range(3):
        for got in (host[k], dev[k]):
            assert torch.equal(got["image"], ref[k]["image"])
            assert torch.allclose(got["instances"].gt_boxes.tensor, O._bt(ref[k]["instances"].gt_boxes), atol=1e-4)
            assert torch.equal(got["instances"].gt_classes, ref[k]["instances"].gt_classes)
            assert got["instances"].image_size == (H, W) and got["file_name"] == f"img{k}"
        assert torch.equal(data[k]["image"], batch[k]["image"])  # inputs untouched (the reference deep-copies, :558)
        assert torch.equal(data[k]["instances"].gt_boxes.tensor, batch[k]["instances"].gt_boxes.tensor)


def test_model_targets_padding_host_logic():
    """`GuassianGeneralizedRCNN._targets`: ground truth / pseudo labels of a batch -> fixed-capacity tensors + counts
    (capacity = next multiple of 16 >= the largest count, >= 16; images without boxes give count 0; device-resident
    pseudo labels keep their device-side count)."""
    import types
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import GuassianGeneralizedRCNN
    me = types.SimpleNamespace(device=torch.device("cpu"))
    g = torch.Generator().manual_seed(0)

    def inst(m):
        xy = torch.rand(m, 2, generator=g) * 50
        return FreeInstances((100, 120), gt_boxes=Boxes(torch.cat([xy, xy + 10], 1)), gt_classes=torch.arange(m) % 8)
    batch = [inst(3), inst(0), inst(17)]
    t = GuassianGeneralizedRCNN._targets(me, batch)
    assert t["gt_boxes"].shape == (3, 32, 4) and t["gt_classes"].shape == (3, 32) and t["gt_classes"].dtype == torch.int32
    assert t["gt_count"].tolist() == [3, 0, 17]
    for k, i in enumerate(batch):
        m = len(i.gt_boxes)
        assert torch.equal(t["gt_boxes"][k, :m], i.gt_boxes.tensor) and float(t["gt_boxes"][k, m:].abs().sum()) == 0
        assert torch.equal(t["gt_classes"][k, :m].long(), i.gt_classes)
    # pseudo labels: capacity 100 buffers with a device-side count (the teacher's output fed back unchanged)
    def pseudo(n_valid):
        p = FreeInstances((100, 120), pseudo_boxes=Boxes(torch.rand(100, 4, generator=g)),
                          scores_logists=torch.rand(100, 9, generator=g), boxes_sigma=torch.rand(100, 4, generator=g))
        p._count = torch.tensor(n_valid)
        return p
    t = GuassianGeneralizedRCNN._targets(me, [pseudo(100), pseudo(37)])
    assert t["pseudo_boxes"].shape == (2, 100, 4) and t["scores_logists"].shape == (2, 100, 9)
    assert t["pseudo_count"].tolist() == [100, 37] and t["pseudo_count"].dtype == torch.int32
    # exact-length pseudo labels of different lengths (e.g. loaded from the reference): padded to the longest
    a = FreeInstances((100, 120), pseudo_boxes=Boxes(torch.rand(5, 4, generator=g)), scores_logists=torch.rand(5, 9, generator=g),
                      boxes_sigma=torch.rand(5, 4, generator=g))
    b = FreeInstances((100, 120), pseudo_boxes=Boxes(torch.zeros(0, 4)), scores_logists=torch.zeros(0, 9), boxes_sigma=torch.zeros(0, 4))
    t = GuassianGeneralizedRCNN._targets(me, [a, b])
    assert t["pseudo_boxes"].shape == (2, 16, 4) and t["pseudo_count"].tolist() == [5, 0]
    assert torch.equal(t["pseudo_boxes"][0, :5], a.pseudo_boxes.tensor) and float(t["pseudo_boxes"][1].abs().sum()) == 0
