"""GPU: rows SURVEY.md 8f marks "next" (checkpoint I/O, eval path) and the burn-in branch of the trainer.

* `PTrainer.save_checkpoint` / `resume_or_load` on the real arenas (`pt/engine/trainer.py:104-111,466-495`,
  `pt/checkpoint/detection_checkpoint.py`): a second trainer resumed from the file holds bit-identical student,
  teacher, momentum and fp16 operand arenas and continues at the next iteration; the file carries the reference's key
  names (`modelTeacher.` / `modelStudent.` + detectron2 parameter names) and layouts. (Host logic:
  tests/test_checkpoint_cpu.py.)
* source-only (burn-in) steps against the reference's own trainer (tests/golden/pt_reference_burnin_golden.pt).
* eval-mode inference + `detector_postprocess` against the reference's own model classes in eval mode
  (tests/golden/pt_reference_eval_golden.pt).
(The file sorts last on purpose: these tests were added after the round's last GPU session.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_trainer_checkpoint_resume(cuda, tmp_path):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    from probabilisticteacher_b200.synthetic import synthetic_batch
    H, W, K = 192, 272, 8
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.OUTPUT_DIR = str(tmp_path)

    def loader():
        s = 0
        while True:
            lab = synthetic_batch(2, H, W, K, 10 + s)
            unl = synthetic_batch(2, H, W, K, 500 + s, labelled=False)
            yield lab, [dict(d) for d in lab], unl, [dict(d) for d in unl]
            s += 1

    tr = PTrainer(cfg, loader(), device=cuda, seed=3)
    for _ in range(2):
        tr.run_step()
    path = tr.save_checkpoint()
    assert path.endswith("model_0000001.pth")
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert raw["iteration"] == 1 and set(raw) == {"model", "optimizer", "scheduler", "iteration"}
    # the file speaks the reference's names and layouts: the oracle (which takes reference state dicts) loads it
    om = O.OracleRCNN(O.OracleCfg(num_classes=K), seed=0)
    student_sd = {k[len("modelStudent."):]: v for k, v in raw["model"].items() if k.startswith("modelStudent.")}
    assert set(student_sd) == set(om.ref_state_dict())
    om.load_ref_state_dict(student_sd)
    for k, v in om.ref_state_dict().items():
        assert v.shape == student_sd[k].shape and torch.equal(v.detach(), student_sd[k]), k

    tr2 = PTrainer(cfg, loader(), device=cuda, seed=99)
    assert not torch.equal(tr2.model.arena.data, tr.model.arena.data)
    start = tr2.resume_or_load(resume=True)
    torch.cuda.synchronize()
    assert start == 2 and tr2.iter == 2
    def same(sd_a, sd_b):
        assert set(sd_a) == set(sd_b)
        for k in sd_a:
            assert torch.equal(sd_a[k], sd_b[k]), k

    for a, b in ((tr.model.arena, tr2.model.arena), (tr.model_teacher.arena, tr2.model_teacher.arena)):
        same(a.state_dict(), b.state_dict())  # every parameter, bit for bit (arena padding is not part of the contract)
        for s in a.segments.values():         # fp16 GEMM operands re-packed from the loaded masters
            if s.kind in ("conv", "fc1", "mat"):
                assert torch.equal(a.hview(s.name), b.hview(s.name)), s.name
    same(tr.model.arena.momentum_state_dict(), tr2.model.arena.momentum_state_dict())
    assert sum(float(v.abs().sum()) for v in tr.model.arena.momentum_state_dict().values()) > 0
    # the teacher differs from the student after an EMA step: the two prefixes did not get mixed up
    k = "roi_heads.box_predictor.cls_score.weight"
    assert not torch.equal(tr2.model_teacher.state_dict()[k], tr2.model.state_dict()[k])
    # and the resumed trainer steps
    losses = tr2.run_step()
    torch.cuda.synchronize()
    assert len(losses) == 8 and all(torch.isfinite(v).all() for v in losses.values())

    # weights-only load (resume=False): momentum and iteration stay at their initial values
    cfg3 = c2f_config()
    cfg3.UNSUPNET.BURN_UP_STEP = 0
    cfg3.OUTPUT_DIR = str(tmp_path / "other")
    cfg3.MODEL.WEIGHTS = path
    tr3 = PTrainer(cfg3, loader(), device=cuda, seed=5)
    assert tr3.resume_or_load(resume=False) == 0
    same(tr3.model.state_dict(), tr.model.state_dict())
    same(tr3.model_teacher.state_dict(), tr.model_teacher.state_dict())
    assert float(tr3.model.arena.momentum.abs().sum()) == 0.0


@pytest.mark.parametrize("precision", ["f16x3", "f16"])
def test_burn_in_steps_vs_reference_trainer(cuda, precision):
    """(precision="f16x3": losses and per-tensor updates within max(1e-3, 4 x the fixture's measured conditioning),
    tests/golden/step_conditioning.json.) Source-only iterations (iter < BURN_UP_STEP, pt/engine/trainer.py:274-290) of the B200 `PTrainer.run_step`
    against the reference's own trainer (tests/golden/pt_reference_burnin_golden.pt, oracle/make_golden_burnin.py):
    same weights, images, `resize` ratios (q views first, then k) and sampling priorities. fp16 operands, so the
    tolerances are those of tests/test_trainer_step_gpu.py; the teacher must stay untouched."""
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_burnin_golden.pt"), weights_only=False)
    H, W, K = G["H"], G["W"], G["K"]
    cfg = c2f_config()
    assert cfg.UNSUPNET.BURN_UP_STEP > 2
    cfg.SOLVER.WARMUP_ITERS = 0
    cfg.SOLVER.BASE_LR = G["lr"]

    def view(tag):
        return [{"image": im.clone(), "height": H, "width": W,
                 "instances": FreeInstances((H, W), gt_boxes=Boxes(b.clone()), gt_classes=c.clone())}
                for im, b, c in zip(G[f"lab_{tag}_images"], G[f"gt_boxes_{tag}"], G[f"gt_classes_{tag}"])]

    def loader():
        while True:
            yield view("q"), view("k"), [], []

    class _Ratios:
        def __init__(self, draws):
            self.draws = list(draws)

        def uniform(self, a, b):
            return self.draws.pop(0)

    def idx(numel, n=64):
        g = torch.Generator().manual_seed(numel)
        return torch.randint(0, numel, (min(n, numel),), generator=g)

    def samples(model):
        return {k: v.detach().reshape(-1).cpu()[idx(v.numel())] for k, v in model.state_dict().items()}

    tr = PTrainer(cfg, loader(), device=cuda, seed=0, precision=precision)
    x3 = precision == "f16x3"
    import json
    COND = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "step_conditioning.json")))["pt_reference_burnin_golden.pt"]
    sd = {k: v.detach() for k, v in O.OracleRCNN(O.OracleCfg(num_classes=K), seed=G["seed"]).ref_state_dict().items()}
    tr.model.load_state_dict(sd)
    tr.model.prio_override = {k: (v[0].to(cuda), v[1].to(cuda)) for k, v in G["prio"].items()}
    teacher_before = tr.model_teacher.arena.data.clone()
    init = {k: v.reshape(-1)[idx(v.numel())] for k, v in sd.items()}
    prev, prev_ref = init, init
    problems = []
    for it, ref in enumerate(G["steps"]):
        tr.rng = _Ratios(ref["ratios"])
        losses = tr.run_step()
        torch.cuda.synchronize()
        assert tr.rng.draws == []  # one ratio per image of q + k
        got = {k: float(v) for k, v in losses.items()}
        assert set(got) == set(ref["losses"])  # un-suffixed keys during burn-in
        print("burn-in step", it, {k: (round(got[k], 4), round(v, 4)) for k, v in ref["losses"].items()})
        if it == 0 or x3:
            for k, v in ref["losses"].items():
                tol = 2e-2 if "rpn" in k else 0.1  # ROI losses depend on the fp16-sensitive proposal selection
                if x3:
                    tol = max(1e-3, 4 * COND[it]["losses"][k])
                    print(f"      {k}: rel err {abs(got[k] - v) / max(abs(v), 1e-3):.2e} (tolerance {tol:.1e})")
                if not abs(got[k] - v) <= tol * max(abs(v), 1e-3):
                    problems.append((it, k, got[k], v, tol))
        st = samples(tr.model)
        if x3:
            worst = ("", 0.0, 0.0)
            for k in sorted(st):
                du, dr = st[k] - prev[k], ref["student"][k] - prev_ref[k]
                if float(dr.abs().max()) > 0.0:
                    e = float((du - dr).abs().max() / dr.abs().max())
                    tol = max(1e-3, 4 * COND[it]["update"].get(k, 0.0))
                    worst = (k, e / tol, e) if e / tol > worst[1] else worst
                    if e > tol:
                        problems.append(("update_x3", it, k, e, tol))
            print("   worst per-tensor update error / tolerance", worst)
        up_g = torch.cat([st[k] - prev[k] for k in sorted(st)])
        up_r = torch.cat([ref["student"][k] - prev_ref[k] for k in sorted(st)])
        cos = float(torch.dot(up_g, up_r) / (up_g.norm() * up_r.norm()))
        ratio = float(up_g.norm() / up_r.norm())
        print("   update cosine", round(cos, 4), "norm ratio", round(ratio, 4))
        if not (cos > (0.99 if it == 0 else 0.98) and (0.95 if it == 0 else 0.9) < ratio < (1.05 if it == 0 else 1.1)):
            problems.append(("update", it, cos, ratio))
        prev, prev_ref = st, ref["student"]
    assert torch.equal(tr.model_teacher.arena.data, teacher_before)
    assert tr.iter == 2
    assert not problems, problems


@pytest.mark.parametrize("case", ["c2f_upscaled", "k1_default_anchors_mixed_sizes"])
def test_eval_mode_vs_reference_model_golden(cuda, case):
    """Eval-mode inference of the CUDA path (f16x3 precision) against the REFERENCE'S OWN MODEL CLASSES in eval mode
    (tests/golden/pt_reference_eval_golden.pt, oracle/make_golden_eval.py): test-time top-k, pseudo-label filter,
    `detector_postprocess` to an output size that differs from the input size. Detections are compared as sets
    (near-tied scores at the synthetic initialisation re-order the lists, see tests/test_parity_x3_gpu.py)."""
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    TOL = 1e-3
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_eval_golden.pt"),
                   weights_only=False)[case]
    cfg = c2f_config()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = G["K"]
    cfg.MODEL.ANCHOR_GENERATOR.NAME = G["anchor_generator"]
    model = build_model(cfg, cuda, precision="f16x3", with_grads=False)
    sd = O.OracleRCNN(O.OracleCfg(num_classes=G["K"], anchor_generator=G["anchor_generator"]), seed=G["seed"]).ref_state_dict()
    model.load_state_dict({k: v.detach() for k, v in sd.items()})
    model.eval()
    batch = [{"image": im, "height": oh, "width": ow} for im, (oh, ow) in zip(G["images"], G["out_sizes"])]
    out = model(batch)
    torch.cuda.synchronize()
    assert len(out) == len(batch)
    for n, (o, ref) in enumerate(zip(out, G["detections"])):
        g = o["instances"]
        assert tuple(g.image_size) == tuple(ref["image_size"])
        assert abs(len(g) - len(ref["scores"])) <= 2, (len(g), len(ref["scores"]))
        scale = float(max(ref["image_size"]))
        gb, gc = g.pred_boxes.tensor.double().cpu(), g.pred_classes.cpu()
        ob, oc = ref["pred_boxes"].double(), ref["pred_classes"]
        d = (gb[:, None, :] - ob[None, :, :]).abs().amax(-1) / scale
        d[gc[:, None] != oc[None, :]] = 1e9
        best, idx = d.min(1)
        ok = best < TOL
        assert float(ok.double().mean()) >= 0.9, (n, float(ok.double().mean()))  # (a near-tie may swap one of ~36)
        for f in ("scores", "scores_logists", "boxes_sigma"):
            a, b = getattr(g, f).double().cpu()[ok], ref[f].double()[idx[ok]]
            err = float(((a - b).abs().reshape(len(a), -1).amax(1) / b.abs().max().clamp_min(1e-30)).max())
            assert err < TOL, (n, f, err)
        # boxes live in the OUTPUT resolution and inside it
        assert float(gb[:, 2].max()) <= ref["image_size"][1] and float(gb[:, 3].max()) <= ref["image_size"][0]


def test_resize_kernel_vs_reference_resize(cuda):
    """SURVEY 8a-18 on its own: `ptb200_resize_paste_u8` (+ the device-geometry variant the CUDA-graph step uses) and
    the box bookkeeping of `PTrainer.resize` / `resize_dev` against the oracle's restatement of
    `pt/engine/trainer.py:557-590` (torch `F.interpolate` bilinear, float -> uint8 truncation, `pixel_mean.int()` canvas),
    which the step fixtures pin to the reference's own `resize`. The kernel evaluates the same expression with its own
    FMA contraction, so a result that lands within ~1e-5 of an integer may truncate to the neighbouring value: a CPU
    emulation of the kernel's arithmetic differs from torch in <= 11 of 3.2 M pixels at 3x800x1333 (never by more than
    1); asserted here: no pixel differs by more than 1 and at most max(8, 1e-4 of them) differ at all."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.engine.trainer import PTrainer
    from probabilisticteacher_b200.structures import Boxes, FreeInstances

    class _Ratios:
        def __init__(self, draws):
            self.draws = list(draws)

        def uniform(self, a, b):
            return self.draws.pop(0)

    mean = torch.tensor([103.53, 116.28, 123.675])
    tr = PTrainer.__new__(PTrainer)  # only the fields `resize` reads: no models are built
    tr.device, tr._pix = cuda, [int(x) for x in mean]
    for (H, W), ratios in (((800, 1333), [0.5, 0.6180339, 0.9731]), ((96, 131), [1.0, 0.75, 0.5000001])):
        n = len(ratios)
        batch = O.synthetic_batch(n, H, W, 8, 5)
        ref = O.resize_batch(batch, ratios, mean)

        def mine(images):
            return [{"image": im, "height": H, "width": W,
                     "instances": FreeInstances((H, W), gt_boxes=Boxes(d["instances"].gt_boxes.tensor.clone().to(im.device)),
                                                gt_classes=d["instances"].gt_classes.clone())}
                    for im, d in zip(images, batch)]

        tr.rng = _Ratios(ratios)
        host = tr.resize(mine([d["image"] for d in batch]))           # CPU images in, geometry as arguments
        params = torch.tensor([[int(H * r), int(W * r), int((W - int(W * r)) / 2), int((H - int(H * r)) / 2)]
                               for r in ratios], dtype=torch.int32, device=cuda)
        ratio_dev = torch.tensor(ratios, dtype=torch.float32, device=cuda)
        dev = tr.resize_dev(mine([d["image"].to(cuda) for d in batch]), params, ratio_dev)  # geometry in device memory
        torch.cuda.synchronize()
        for k in range(n):
            want = ref[k]["image"].to(torch.int32)
            for got in (host[k], dev[k]):
                diff = (got["image"].cpu().to(torch.int32) - want).abs()
                n_diff = int((diff > 0).sum())
                assert int(diff.max()) <= 1 and n_diff <= max(8, 1e-4 * diff.numel()), (H, W, ratios[k], int(diff.max()), n_diff)
                d_h, d_w = int(H * ratios[k]), int(W * ratios[k])
                x1, y1 = int((W - d_w) / 2), int((H - d_h) / 2)
                canvas = torch.ones(H, W, dtype=torch.bool)
                canvas[y1:y1 + d_h, x1:x1 + d_w] = False
                for c in range(3):  # outside the paste: exactly int(pixel_mean)
                    assert bool((got["image"][c].cpu()[canvas] == int(mean[c])).all())
                assert torch.allclose(got["instances"].gt_boxes.tensor.cpu(), O._bt(ref[k]["instances"].gt_boxes), atol=1e-3)
                assert torch.equal(got["instances"].gt_classes.cpu(), ref[k]["instances"].gt_classes)
        # inputs are not modified (the reference deep-copies first, trainer.py:558)
        assert torch.equal(batch[0]["image"], O.synthetic_batch(n, H, W, 8, 5)[0]["image"])


def test_anchor_generators_bit_exact(cuda):
    """SURVEY 8a-4 on its own: `ptb200_cell_anchors_from_wh` + `ptb200_anchor_grid` against the oracle's restatement of
    `pt/modeling/anchor_generator.py:108-164` / detectron2's DefaultAnchorGenerator (pinned by the function-level
    fixture and detectron2's known-answer table): every anchor coordinate is one fp32 addition of an exactly
    representable grid shift and a cell coordinate, so the comparison is BIT-EXACT, at the four feature-map sizes of
    BASELINE.json's configs and with a half-stride offset."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.arena import ParamArena
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.anchor_generator import DefaultAnchorGenerator, DifferentiableAnchorGenerator
    cfg = c2f_config()
    arena = ParamArena(num_classes=8, differentiable_anchors=True, device=cuda, with_grads=False)
    # the constructor initialises the learnable (w, h) pairs from cfg (anchor_generator.py:66-72) ...
    custom = [[w * 1.5, h * 0.75] for w, h in cfg.MODEL.ANCHOR_GENERATOR.ANCHOR[0]]
    cfg.MODEL.ANCHOR_GENERATOR.ANCHOR = [custom]
    gen0 = DifferentiableAnchorGenerator(cfg, arena)
    assert torch.equal(arena.view("proposal_generator.anchor_generator.anchor_0").cpu(), torch.tensor(custom))
    assert torch.equal(gen0(3, 5).cpu(), O.grid_anchors(O.differentiable_cell_anchors(torch.tensor(custom)), 3, 5, 16,
                                                        cfg.MODEL.ANCHOR_GENERATOR.OFFSET))
    cfg = c2f_config()
    wh = torch.tensor(cfg.MODEL.ANCHOR_GENERATOR.ANCHOR[0]) * torch.linspace(0.9, 1.13, 18).view(9, 2)  # "learned" values
    for offset in (0.0, 0.5):
        cfg.MODEL.ANCHOR_GENERATOR.OFFSET = offset
        dgen = DifferentiableAnchorGenerator(cfg, arena)
        arena.view("proposal_generator.anchor_generator.anchor_0").copy_(wh)   # ... and training moves them
        gens = ((dgen, O.differentiable_cell_anchors(wh)),
                (DefaultAnchorGenerator(cfg, arena), O.default_cell_anchors(cfg.MODEL.ANCHOR_GENERATOR.SIZES[0],
                                                                            cfg.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS[0])))
        for gen, cell in gens:
            for H, W in ((50, 83), (37, 75), (37, 125), (64, 128), (1, 1)):
                got = gen(H, W)
                torch.cuda.synchronize()
                want = O.grid_anchors(cell, H, W, 16, offset)
                assert got.shape == (H * W * 9, 4)
                assert torch.equal(got.cpu(), want), (type(gen).__name__, H, W, offset)
    # the layout contract of SURVEY 8a-4: anchors[(y*W + x)*9 + a] = (x*16, y*16, x*16, y*16) + cell[a]
    cfg.MODEL.ANCHOR_GENERATOR.OFFSET = 0.0
    a = DefaultAnchorGenerator(cfg, arena)(50, 83).cpu()
    cell = O.default_cell_anchors((128, 256, 512), (0.5, 1.0, 2.0))
    y, x, k = 17, 44, 5
    assert torch.equal(a[(y * 83 + x) * 9 + k], torch.tensor([x * 16., y * 16., x * 16., y * 16.]) + cell[k])


@pytest.mark.parametrize("thr,classes", [(0.5, 0), (0.7, 0), (0.5, 8)])
def test_nms_collisions_bit_exact(cuda, thr, classes):
    """NMS keep indices on the inputs where implementations usually part ways: boxes on an integer grid (many IoUs
    EXACTLY equal to the threshold: `>` must stay strict), exact duplicates, zero-area boxes and heavily tied scores
    (order = stable sort, as torchvision / the oracle). The oracle agrees with torchvision on this family
    (tests/test_oracle_cpu.py::test_nms_ties_duplicates_and_exact_threshold)."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200 import ops
    g = torch.Generator().manual_seed(int(thr * 10) + classes)
    n = 3000
    xy = torch.randint(0, 40, (n, 2), generator=g).float() * 4
    wh = torch.randint(0, 12, (n, 2), generator=g).float() * 8   # includes zero-width / zero-height boxes
    b = torch.cat([xy, xy + wh], 1)
    b[::7] = b[0].clone()                                         # exact duplicates
    s = torch.randint(0, 5, (n,), generator=g).float() / 4       # five distinct scores
    order = torch.argsort(-s, stable=True).to(torch.int32)
    if classes:
        ref = O.batched_nms(b, s, torch.arange(n) % classes, thr)
    else:
        ref = O.nms(b, s, thr)
    max_keep = 2000
    cap = ((n + 63) // 64) * 64
    order_p = torch.zeros(1, cap, dtype=torch.int32)
    order_p[0, :n] = order
    keep_idx, keep_count = ops.nms(b[None].to(cuda), order_p.to(cuda), torch.tensor([n], dtype=torch.int32).to(cuda),
                                   thr, max_keep, class_mod=classes)
    torch.cuda.synchronize()
    kc = int(keep_count[0])
    got = order[keep_idx[0, :kc].cpu().long()].long()
    assert kc == min(len(ref), max_keep)
    assert torch.equal(got, ref[:max_keep])


def test_rpn_match_collisions(cuda):
    """Anchor labelling on the inputs where the Matcher's tie rules decide (detectron2 `Matcher`): a
    DUPLICATED ground-truth box (arg-max tie -> the first index), a ZERO-AREA ground-truth box (its best IoU is 0, so
    the low-quality rule promotes every anchor whose IoU with it is 0 -- detectron2's quirk, kept), a box that
    coincides exactly with an anchor (IoU 1) and an image with a single far-off-grid box. Indices and labels must equal
    the oracle's bit for bit."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.modeling import sampling
    H, W = 25, 38
    cell = O.default_cell_anchors((64, 128, 256), (0.5, 1.0, 2.0))
    anchors = O.grid_anchors(cell, H, W, 16, 0.0)
    R = anchors.shape[0]
    N, cap = 3, 64
    gt = torch.zeros(N, cap, 4)
    a_mid = anchors[(12 * W + 20) * 9 + 4].clone()          # exactly an anchor
    normal = torch.tensor([[40., 60., 200., 180.], [300., 100., 420., 330.], [100., 200., 180., 390.]])
    rows = [torch.cat([normal, normal[:1], a_mid[None]]),                        # duplicate of box 0 + exact anchor
            torch.cat([normal[:2], torch.tensor([[250., 250., 250., 300.]])]),   # zero-area box
            torch.tensor([[3., 5., 9., 11.]])]                                   # one tiny box near the corner
    cnt = torch.tensor([len(r) for r in rows], dtype=torch.int32)
    for n, r in enumerate(rows):
        gt[n, :len(r)] = r
    matched, labels = sampling.rpn_match(gt.to(cuda), cnt.to(cuda), anchors.to(cuda), N, 0.3, 0.7)
    torch.cuda.synchronize()
    for n in range(N):
        iou = O.pairwise_iou(gt[n, :cnt[n]], anchors)
        rm, rl = O.matcher(iou, (0.3, 0.7), (0, -1, 1), True)
        assert torch.equal(labels[n].cpu().to(torch.int8), rl), n
        assert torch.equal(matched[n].cpu().long(), rm), n
    # what the cases are meant to exercise really happens
    iou0 = O.pairwise_iou(gt[0, :cnt[0]], anchors)
    assert torch.equal(iou0[0], iou0[3]) and int((matched[0].cpu() == 3).sum()) == 0      # tie -> first index
    assert float(iou0[4].max()) == 1.0
    lab1 = labels[1].cpu()
    iou1 = O.pairwise_iou(gt[1, :cnt[1]], anchors)
    assert float(iou1[2].max()) == 0.0 and int((lab1 == 1).sum()) >= int((iou1.max(0).values == 0).sum()) > 0


def test_nonfinite_proposals_are_dropped_and_flagged(cuda):
    """Error convention of SURVEY 8b: the reference raises FloatingPointError when a decoded proposal or its score is
    Inf/NaN in training and silently filters such rows in eval (`proposal_utils.py:117-127`). Here the decode kernel
    always filters them and sets a device flag that `GuassianRPN.raise_if_nonfinite` turns into the reference's
    error lazily. With one NaN logit and one Inf delta injected, the selected proposals must equal the oracle's
    eval-mode result on the same inputs (the two bad anchors gone, everything else bit-identical in order)."""
    import types
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.modeling.proposal_generator.proposal_utils import find_top_rpn_proposals
    from probabilisticteacher_b200.modeling.proposal_generator.rpn import GuassianRPN
    g = torch.Generator().manual_seed(12)
    N, H, W, A = 2, 12, 17, 9
    R = H * W * A
    anchors = O.grid_anchors(O.differentiable_cell_anchors(torch.tensor(O.OracleCfg().anchor_wh)), H, W, 16, 0.0)
    lg = torch.zeros(N, H, W + 1, A)
    lg[:, :, :W] = torch.randn(N, H, W, A, generator=g)
    dl = torch.zeros(N, H, W + 1, A * 8)
    dl[:, :, :W] = torch.randn(N, H, W, A * 8, generator=g) * 0.2
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    hw = torch.tensor([[H * 16.0, W * 16.0]] * N, device=cuda)

    def run():
        out = find_top_rpn_proposals(lg.reshape(N, -1, A).to(cuda), dl.reshape(N, -1, A * 8).to(cuda), anchors.to(cuda),
                                     N, H, W, A, hw, 0.7, 12000, 2000, 0.0, flag)
        torch.cuda.synchronize()
        return out

    run()
    rpn = types.SimpleNamespace(nonfinite_flag=flag)
    GuassianRPN.raise_if_nonfinite(rpn)  # clean inputs: no error
    lg[0, 3, 5, 2] = float("nan")
    dl[1, 7, 9, 4 * 8 + 0] = float("inf")  # dx of anchor 4 at (7, 9): the decoded box is non-finite (dw / dh are clamped)
    boxes, scores, count = run()
    assert int(flag) != 0
    with pytest.raises(FloatingPointError):
        GuassianRPN.raise_if_nonfinite(rpn)
    assert int(flag) == 0  # cleared by the check
    # oracle, eval-mode filtering (training=False keeps the reference from raising); same top-k sizes as the call above
    lo = lg[:, :, :W].reshape(N, R)
    do = dl[:, :, :W].reshape(N, R, 8)
    pr = O.apply_deltas(do[..., :4].reshape(-1, 4), anchors.expand(N, R, 4).reshape(-1, 4), (1., 1., 1., 1.)).view(N, R, 4)
    ref = O.find_top_rpn_proposals(pr, lo.clone(), [(H * 16, W * 16)] * N, 0.7, 12000, 2000, 0, False, do[..., 4:])
    for n in range(N):
        c = int(count[n])
        rb, rs = O._bt(ref[n].proposal_boxes), ref[n].objectness_logits
        assert c == len(rs), (n, c, len(rs))
        assert (boxes[n, :c].cpu() - rb).abs().max() < 1e-2 and (scores[n, :c].cpu() - rs).abs().max() < 1e-5
        assert bool(torch.isfinite(boxes[n, :c]).all()) and bool(torch.isfinite(scores[n, :c]).all())


@pytest.mark.parametrize("case", ["config1", "config4"])
def test_full_size_configs_vs_reference_model_golden(cuda, case):
    """BASELINE.json config 1 and config 4's model / size at FULL SIZE against the REFERENCE'S OWN MODEL CLASSES
    (tests/golden/pt_reference_{case}_golden.pt, oracle/make_golden_config1.py: Guassian-RCNN-VGG.yaml's model, 1 source
    + 1 target synthetic 3x800x1333 image; final_k2c.yaml's K = 1 model with differentiable anchors at 3x600x2000; the
    forward passes of one post-burn-in iteration). CUDA path in the f16x3
    parity precision, every stage selecting its OWN proposals. Asserted at 1e-3: the RPN losses of both student
    branches (they do not depend on proposal selection); asserted as sets: the teacher's proposals and pseudo labels.
    The ROI-stage losses are printed only: at full size a near-threshold NMS flip changes which rois get sampled
    (tests/test_parity_x3_gpu.py::_full_iteration compares them on shared proposals against the oracle, which
    reproduces this fixture exactly: tests/test_oracle_golden_model.py)."""
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    TOL = 1e-3
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", f"pt_reference_{case}_golden.pt"), weights_only=False)
    H, W, K, N = G["H"], G["W"], G["K"], G["N"]
    cfg = c2f_config()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = K
    cfg.MODEL.ANCHOR_GENERATOR.NAME = G["anchor_generator"]
    model = build_model(cfg, cuda, precision="f16x3", with_grads=False)
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K, anchor_generator=G["anchor_generator"]), seed=G["weight_seed"]).ref_state_dict()
    model.load_state_dict({k: v.detach() for k, v in sd.items()})
    model.train()
    g = torch.Generator().manual_seed(G["prio_seed"])
    R, L = (H // 16) * (W // 16) * 9, 2000 + 16
    model.prio_override = {"rpn": (torch.rand(N, R, generator=g).to(cuda), torch.rand(N, R, generator=g).to(cuda)),
                           "roi": (torch.rand(N, L, generator=g).to(cuda), torch.rand(N, L, generator=g).to(cuda))}
    lab = [{"image": d["image"], "height": H, "width": W,
            "instances": FreeInstances((H, W), gt_boxes=Boxes(d["instances"].gt_boxes.tensor.clone()),
                                       gt_classes=d["instances"].gt_classes.clone())}
           for d in O.synthetic_batch(N, H, W, K, G["lab_seed"])]
    unl = [{"image": d["image"], "height": H, "width": W} for d in O.synthetic_batch(N, H, W, K, G["unl_seed"], labelled=False)]
    scale = float(max(H, W))
    with torch.no_grad():
        ls, _, _, _ = model(lab, branch="supervised")
        print("sup", {k: (round(float(ls[k]), 5), round(v, 5)) for k, v in G["sup_losses"].items()})
        for k in ("loss_rpn_cls", "loss_rpn_loc"):
            a, b = float(ls[k]), G["sup_losses"][k]
            assert abs(a - b) <= TOL * max(abs(b), 1e-6), ("sup", k, a, b)
        _, pg, rg, _ = model(unl, branch="unsup_data_weak")
        for n in range(N):
            p = pg[n].trim()
            ref_boxes = G["teacher_rpn_boxes"][n]
            assert abs(len(p) - len(ref_boxes)) <= max(2, len(ref_boxes) // 50), (len(p), len(ref_boxes))
            a, b = p.proposal_boxes.tensor.double().cpu(), ref_boxes.double()
            d = (a[:, None, :] - b[None, :, :]).abs().amax(-1) / scale
            assert float((d.min(1).values < TOL).double().mean()) >= 0.95
            ref = G["teacher_roih"][n]
            det = rg[n].trim()
            assert abs(len(det) - len(ref["scores"])) <= 2
            gb, gc = det.pred_boxes.tensor.double().cpu(), det.pred_classes.cpu()
            dd = (gb[:, None, :] - ref["pred_boxes"].double()[None, :, :]).abs().amax(-1) / scale
            dd[gc[:, None] != ref["pred_classes"][None, :]] = 1e9
            frac = float((dd.min(1).values < TOL).double().mean())
            print("teacher", len(p), len(ref_boxes), "detections matched", frac)
            assert frac >= 0.9, frac
        unl_q = [dict(d, instances=FreeInstances((H, W), pseudo_boxes=Boxes(r["pred_boxes"].to(cuda)),
                                                 scores_logists=r["scores_logists"].to(cuda),
                                                 boxes_sigma=r["boxes_sigma"].to(cuda)))
                 for d, r in zip(unl, G["teacher_roih"])]
        lu, _, _, _ = model(unl_q, branch="unsupervised", danchor=True)
        print("unsup", {k: (round(float(lu[k]), 5), round(v, 5)) for k, v in G["unsup_losses"].items()})
        for k in ("loss_rpn_cls", "loss_rpn_loc"):
            a, b = float(lu[k]), G["unsup_losses"][k]
            assert abs(a - b) <= TOL * max(abs(b), 1e-6), ("unsup", k, a, b)


@pytest.mark.parametrize("case", ["second_image_empty", "all_empty"])
def test_unsupervised_branch_without_pseudo_labels(cuda, case):
    """The "empty input" edge of the unsupervised branch against the reference's own model classes
    (tests/golden/pt_reference_empty_pseudo_golden.pt, oracle/make_golden_empty_pseudo.py): no pseudo label for one
    image -> all four losses at 1e-3 (f16x3); none at all -> both RPN losses exactly 0, and the ROI losses, which the
    reference returns as NaN (a mean over zero rois), are NaN or 0 here (a guarded division is acceptable: training has
    nothing to learn from such a batch either way)."""
    import math
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_empty_pseudo_golden.pt"), weights_only=False)
    H, W, K, N = G["H"], G["W"], G["K"], G["N"]
    model = build_model(c2f_config(), cuda, precision="f16x3", with_grads=False)
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K), seed=G["weight_seed"]).ref_state_dict()
    model.load_state_dict({k: v.detach() for k, v in sd.items()})
    model.train()
    g = torch.Generator().manual_seed(G["prio_seed"])
    R, L = (H // 16) * (W // 16) * 9, 2000 + 16
    model.prio_override = {"rpn": (torch.rand(N, R, generator=g).to(cuda), torch.rand(N, R, generator=g).to(cuda)),
                           "roi": (torch.rand(N, L, generator=g).to(cuda), torch.rand(N, L, generator=g).to(cuda))}
    unl = O.synthetic_batch(N, H, W, K, G["unl_seed"], labelled=False)
    c = G["cases"][case]
    q = []
    for d, r, n in zip(unl, G["teacher_roih"], c["keep"]):
        n = len(r["pred_boxes"]) if n is None else n
        q.append({"image": d["image"], "height": H, "width": W,
                  "instances": FreeInstances((H, W), pseudo_boxes=Boxes(r["pred_boxes"][:n].to(cuda)),
                                             scores_logists=r["scores_logists"][:n].to(cuda),
                                             boxes_sigma=r["boxes_sigma"][:n].to(cuda))})
    with torch.no_grad():
        lu, _, _, _ = model(q, branch="unsupervised", danchor=True)
    torch.cuda.synchronize()
    got = {k: float(v) for k, v in lu.items()}
    print(case, got, c["losses"])
    for k, v in c["losses"].items():
        if math.isnan(v):
            assert math.isnan(got[k]) or got[k] == 0.0, (k, got[k])
        elif v == 0.0:
            assert got[k] == 0.0, (k, got[k])
        else:
            assert abs(got[k] - v) <= 1e-3 * abs(v), (k, got[k], v)


def test_supervised_branch_without_any_ground_truth(cuda):
    """A supervised batch in which NO image has a ground-truth box, against the reference's own model classes
    (tests/golden/pt_reference_empty_pseudo_golden.pt): classification losses at 1e-3 (f16x3), both regression losses
    zero."""
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_empty_pseudo_golden.pt"), weights_only=False)
    H, W, K, N = G["H"], G["W"], G["K"], G["N"]
    model = build_model(c2f_config(), cuda, precision="f16x3", with_grads=False)
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K), seed=G["weight_seed"]).ref_state_dict()
    model.load_state_dict({k: v.detach() for k, v in sd.items()})
    model.train()
    g = torch.Generator().manual_seed(G["prio_seed"])
    R, L = (H // 16) * (W // 16) * 9, 2000 + 16
    model.prio_override = {"rpn": (torch.rand(N, R, generator=g).to(cuda), torch.rand(N, R, generator=g).to(cuda)),
                           "roi": (torch.rand(N, L, generator=g).to(cuda), torch.rand(N, L, generator=g).to(cuda))}
    c = G["supervised_no_gt"]
    lab = [{"image": d["image"], "height": H, "width": W,
            "instances": FreeInstances((H, W), gt_boxes=Boxes(torch.zeros(0, 4)), gt_classes=torch.zeros(0, dtype=torch.int64))}
           for d in O.synthetic_batch(N, H, W, K, c["lab_seed"], boxes_per_image=0)]
    with torch.no_grad():
        ls, _, _, _ = model(lab, branch="supervised")
    torch.cuda.synchronize()
    got = {k: float(v) for k, v in ls.items()}
    print(got, c["losses"])
    for k, v in c["losses"].items():
        if v == 0.0:
            assert abs(got[k]) <= 1e-12, (k, got[k])
        else:
            assert abs(got[k] - v) <= 1e-3 * abs(v), (k, got[k], v)


@pytest.mark.parametrize("variant", [0, 1])
def test_unsupervised_branch_other_unsupnet_settings(cuda, variant):
    """The `efl` / `tau` / `lambda` arguments of the unsupervised loss kernels away from train.sh's values (EFL off with
    TAU [0.25, 0.25], the value configs/pt/final_c2f.yaml itself carries; EFL_LAMBDA [1, 2] with TAU [0.5, 0.25]) against
    the reference's own model classes (tests/golden/pt_reference_empty_pseudo_golden.pt): all four losses of the
    unsupervised branch at 1e-3 in the f16x3 precision."""
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_empty_pseudo_golden.pt"), weights_only=False)
    v = G["unsupnet_variants"][variant]
    H, W, K, N = G["H"], G["W"], G["K"], G["N"]
    cfg = c2f_config()
    cfg.UNSUPNET.EFL, cfg.UNSUPNET.TAU, cfg.UNSUPNET.EFL_LAMBDA = v["efl"], list(v["tau"]), list(v["efl_lambda"])
    model = build_model(cfg, cuda, precision="f16x3", with_grads=False)
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K), seed=G["weight_seed"]).ref_state_dict()
    model.load_state_dict({k: t.detach() for k, t in sd.items()})
    model.train()
    g = torch.Generator().manual_seed(G["prio_seed"])
    R, L = (H // 16) * (W // 16) * 9, 2000 + 16
    model.prio_override = {"rpn": (torch.rand(N, R, generator=g).to(cuda), torch.rand(N, R, generator=g).to(cuda)),
                           "roi": (torch.rand(N, L, generator=g).to(cuda), torch.rand(N, L, generator=g).to(cuda))}
    unl = O.synthetic_batch(N, H, W, K, G["unl_seed"], labelled=False)
    q = [{"image": d["image"], "height": H, "width": W,
          "instances": FreeInstances((H, W), pseudo_boxes=Boxes(r["pred_boxes"].to(cuda)),
                                     scores_logists=r["scores_logists"].to(cuda), boxes_sigma=r["boxes_sigma"].to(cuda))}
         for d, r in zip(unl, G["teacher_roih"])]
    with torch.no_grad():
        lu, _, _, _ = model(q, branch="unsupervised", danchor=True)
    torch.cuda.synchronize()
    got = {k: float(t) for k, t in lu.items()}
    print(v["efl"], v["tau"], v["efl_lambda"], got, v["losses"])
    for k, ref in v["losses"].items():
        assert abs(got[k] - ref) <= 1e-3 * abs(ref), (k, got[k], ref)


def test_every_hyper_parameter_away_from_its_default(cuda):
    """The CUDA path (f16x3) under a configuration in which EVERY honoured hyper-parameter differs from the defaults
    (pixel std, anchor offset, RPN / ROI IoU thresholds, batch sizes, fractions, top-k sizes 300 / 50, pseudo-label filter
    thresholds, 20 detections per image, box-regression weights), against the reference's own model classes
    (tests/golden/pt_reference_oddcfg_golden.pt, oracle/make_golden_oddcfg.py): a kernel or wrapper that ignored one of
    them cannot reproduce these losses."""
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    TOL = 1e-3
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_oddcfg_golden.pt"), weights_only=False)
    H, W, K, N = G["H"], G["W"], G["K"], G["N"]
    cfg = c2f_config()
    for key, value in G["overrides"]:
        node = cfg
        *parents, leaf = key.split(".")
        for p in parents:
            node = node[p]
        node[leaf] = value
    model = build_model(cfg, cuda, precision="f16x3", with_grads=False)
    sd = O.OracleRCNN(O.OracleCfg(num_classes=K, **G["oracle_kw"]), seed=G["weight_seed"]).ref_state_dict()
    model.load_state_dict({k: t.detach() for k, t in sd.items()})
    model.train()
    g = torch.Generator().manual_seed(G["prio_seed"])
    R, L = (H // 16) * (W // 16) * 9, G["roi_prio_len"]
    model.prio_override = {"rpn": (torch.rand(N, R, generator=g).to(cuda), torch.rand(N, R, generator=g).to(cuda)),
                           "roi": (torch.rand(N, L, generator=g).to(cuda), torch.rand(N, L, generator=g).to(cuda))}
    lab = [{"image": d["image"], "height": H, "width": W,
            "instances": FreeInstances((H, W), gt_boxes=Boxes(d["instances"].gt_boxes.tensor.clone()),
                                       gt_classes=d["instances"].gt_classes.clone())}
           for d in O.synthetic_batch(N, H, W, K, G["lab_seed"], boxes_per_image=4)]
    unl = [{"image": d["image"], "height": H, "width": W} for d in O.synthetic_batch(N, H, W, K, G["unl_seed"], labelled=False)]
    scale = float(max(H, W))
    with torch.no_grad():
        ls, _, _, _ = model(lab, branch="supervised")
        print("sup", {k: (round(float(ls[k]), 5), round(v, 5)) for k, v in G["sup_losses"].items()})
        for k, v in G["sup_losses"].items():
            assert abs(float(ls[k]) - v) <= TOL * abs(v), ("sup", k, float(ls[k]), v)
        _, pg, rg, _ = model(unl, branch="unsup_data_weak")
        for n in range(N):
            p, ref_boxes = pg[n].trim(), G["teacher_rpn_boxes"][n]
            assert len(p) == len(ref_boxes), (len(p), len(ref_boxes))
            d = (p.proposal_boxes.tensor.double().cpu()[:, None, :] - ref_boxes.double()[None, :, :]).abs().amax(-1) / scale
            assert float((d.min(1).values < TOL).double().mean()) >= 0.9
            det, ref = rg[n].trim(), G["teacher_roih"][n]
            assert len(det) == len(ref["scores"]) == 20
            dd = (det.pred_boxes.tensor.double().cpu()[:, None, :] - ref["pred_boxes"].double()[None, :, :]).abs().amax(-1) / scale
            dd[det.pred_classes.cpu()[:, None] != ref["pred_classes"][None, :]] = 1e9
            assert float((dd.min(1).values < TOL).double().mean()) >= 0.9
        q = [dict(d, instances=FreeInstances((H, W), pseudo_boxes=Boxes(r["pred_boxes"].to(cuda)),
                                             scores_logists=r["scores_logists"].to(cuda), boxes_sigma=r["boxes_sigma"].to(cuda)))
             for d, r in zip(unl, G["teacher_roih"])]
        lu, _, _, _ = model(q, branch="unsupervised", danchor=True)
        print("unsup", {k: (round(float(lu[k]), 5), round(v, 5)) for k, v in G["unsup_losses"].items()})
        for k, v in G["unsup_losses"].items():
            assert abs(float(lu[k]) - v) <= TOL * abs(v), ("unsup", k, float(lu[k]), v)


@pytest.mark.parametrize("precision", ["f16x3", "f16"])
def test_trainer_hyper_parameters_off_their_defaults(cuda, precision):
    """(precision="f16x3": every loss of every step, the student's per-tensor updates and the teacher's EMA are held
    to north_star's 1e-3 against the reference's own trainer. precision="f16": fp16-operand tolerances below; the EMA
    kernel is then checked exactly against THIS trainer's own student, and against the reference within the share
    (1 - keep) of the student's deviation that the teacher inherits -- round 1 compared the fp16-mode teacher with the
    reference at 2e-4 of max|w| and failed on hardware by 3.7e-6 vs 1.6e-6 on bbox_pred.weight, whose weights are
    ~1e-3: that was the student's fp16-gradient deviation, not the EMA.)
    Three post-burn-in steps of the B200 `PTrainer` with the trainer-level hyper-parameters away from their defaults
    (loss weights 0.5 / 2.0, EMA keep rate 0.99, TEACHER_UPDATE_ITER 2, momentum 0.8, weight decay 5e-4, lr 0.004)
    against the reference's own trainer (tests/golden/pt_reference_step_oddcfg_golden.pt,
    oracle/make_golden_step_oddcfg.py). Exact: the teacher is the initial student after step 0 and UNCHANGED after
    step 1 (TEACHER_UPDATE_ITER); tight: the EMA at step 2 (keep rate); fp16-path tolerances (as in
    tests/test_trainer_step_gpu.py): the losses of step 0 and direction / size of the student's updates (lr, momentum,
    weight decay, clip, loss weights)."""
    import os
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_step_oddcfg_golden.pt"), weights_only=False)
    H, W, K, t = G["H"], G["W"], G["K"], G["trainer_cfg"]
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.UNSUPNET.SOURCE_LOSS_WEIGHT, cfg.UNSUPNET.TARGET_UNSUP_LOSS_WEIGHT = t["source_loss_weight"], t["target_unsup_loss_weight"]
    cfg.UNSUPNET.EMA_KEEP_RATE, cfg.UNSUPNET.TEACHER_UPDATE_ITER = t["ema_keep_rate"], t["teacher_update_iter"]
    cfg.SOLVER.MOMENTUM, cfg.SOLVER.WEIGHT_DECAY, cfg.SOLVER.BASE_LR, cfg.SOLVER.WARMUP_ITERS = \
        t["momentum"], t["weight_decay"], t["base_lr"], 0

    def batch():
        lab = [{"image": im.clone(), "height": H, "width": W,
                "instances": FreeInstances((H, W), gt_boxes=Boxes(b.clone()), gt_classes=c.clone())}
               for im, b, c in zip(G["lab_images"], G["gt_boxes"], G["gt_classes"])]
        unl = [{"image": im.clone(), "height": H, "width": W} for im in G["unl_images"]]
        return lab, unl

    def loader():
        while True:
            lab, unl = batch()
            lab_k, _ = batch()
            _, unl_k = batch()
            yield lab, lab_k, unl, unl_k

    class _Ratios:
        def __init__(self, draws):
            self.draws = list(draws)

        def uniform(self, a, b):
            return self.draws.pop(0)

    def idx(numel, n=64):
        g = torch.Generator().manual_seed(numel)
        return torch.randint(0, numel, (min(n, numel),), generator=g)

    def samples(model):
        return {k: v.detach().reshape(-1).cpu()[idx(v.numel())] for k, v in model.state_dict().items()}

    tr = PTrainer(cfg, loader(), device=cuda, seed=0, precision=precision)
    x3 = precision == "f16x3"
    keep = t["ema_keep_rate"]
    import json
    # conditioning of every fixture quantity under a 2e-6 weight perturbation (oracle/measure_conditioning.py): the
    # f16x3 step is held to max(1e-3, 4 x conditioning), see tests/test_trainer_step_gpu.py
    COND = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "step_conditioning.json")))["pt_reference_step_oddcfg_golden.pt"]
    ocfg = O.OracleCfg(num_classes=K)
    sd = {k: v.detach() for k, v in O.OracleRCNN(ocfg, seed=G["seed"]).ref_state_dict().items()}
    sd_t = {k: v.detach() for k, v in O.OracleRCNN(ocfg, seed=G["teacher_seed"]).ref_state_dict().items()}
    tr.model.load_state_dict(sd)
    tr.model_teacher.load_state_dict(sd_t)
    tr.model.prio_override = {k: (v[0].to(cuda), v[1].to(cuda)) for k, v in G["prio"].items()}
    init = {k: v.reshape(-1)[idx(v.numel())] for k, v in sd.items()}
    prev, prev_ref, teacher_prev = init, init, None
    problems = []
    for it, ref in enumerate(G["steps"]):
        tr.rng = _Ratios(ref["ratios"])
        losses = tr.run_step()
        torch.cuda.synchronize()
        got = {k: float(v) for k, v in losses.items()}
        print("step", it, {k: (round(got[k], 4), round(v, 4)) for k, v in ref["losses"].items()})
        if it == 0 or x3:  # the trainer reports UNWEIGHTED losses (metrics_dict = record_dict, trainer.py:379-381)
            for k, v in ref["losses"].items():
                tol = max(1e-3, 4 * COND[it]["losses"][k]) if x3 else ((2e-2 if "rpn" in k else 0.1) if k.endswith("_sup") else 0.3)
                if not abs(got[k] - v) <= tol * max(abs(v), 1e-3):
                    problems.append((it, k, got[k], v))
        st, te = samples(tr.model), samples(tr.model_teacher)
        if it == 0:
            for k, v in init.items():
                assert torch.equal(te[k], v), k                       # copy of the initial student
        elif it == 1:
            for k in te:
                assert torch.equal(te[k], teacher_prev[k]), k          # TEACHER_UPDATE_ITER = 2: untouched
        else:
            for k, v in ref["teacher"].items():                        # EMA with keep rate 0.99
                own = keep * teacher_prev[k].double() + (1.0 - keep) * prev[k].double()   # prev = own student before this step
                err_own = float((te[k].double() - own).abs().max())
                assert err_own <= 1e-9 + 2e-7 * float(v.abs().max()), (k, err_own)        # the EMA kernel itself (fp32)
                dev = float((prev[k] - prev_ref[k]).abs().max())                          # own student vs reference student
                err = float((te[k] - v).abs().max())
                assert err <= 1e-7 + 2e-6 * float(v.abs().max()) + (1.0 - keep) * dev, (k, err, dev)
                if x3:  # the EMA's change (1 - keep) * (student - teacher), as well-conditioned as the student's updates
                    c = max([COND[j]["update"].get(k, 0.0) for j in range(it)] + [0.0])
                    assert err <= 1e-7 + max(1e-3, 4 * c) * float((v - teacher_prev[k]).abs().max()), (k, err, c)
            assert any(not torch.equal(te[k], teacher_prev[k]) for k in te)
        teacher_prev = te
        up_g = torch.cat([st[k] - prev[k] for k in sorted(st)])
        up_r = torch.cat([ref["student"][k] - prev_ref[k] for k in sorted(st)])
        cos = float(torch.dot(up_g, up_r) / (up_g.norm() * up_r.norm()))
        ratio = float(up_g.norm() / up_r.norm())
        print("   update cosine", round(cos, 6), "norm ratio", round(ratio, 6))
        if x3:
            worst = ("", 0.0, 0.0)
            for k in sorted(st):
                du, dr = st[k] - prev[k], ref["student"][k] - prev_ref[k]
                if float(dr.abs().max()) > 0.0:
                    e = float((du - dr).abs().max() / dr.abs().max())
                    tol = max(1e-3, 4 * COND[it]["update"].get(k, 0.0))
                    worst = (k, e / tol, e) if e / tol > worst[1] else worst
                    if e > tol:
                        problems.append(("update_x3", it, k, e, tol))
            print("   worst per-tensor update error / tolerance", worst)
        elif not (cos > (0.99 if it == 0 else 0.97) and (0.95 if it == 0 else 0.9) < ratio < (1.05 if it == 0 else 1.1)):
            problems.append(("update", it, cos, ratio))
        prev, prev_ref = st, ref["student"]
    assert not problems, problems


def test_eval_mode_postprocess_scaling(cuda):
    """`detector_postprocess` on the device-resident detections of an eval-mode forward: asking for twice the
    resolution doubles every box (exact in fp32) and changes nothing else. Applied twice to the SAME raw detections (a
    second forward may differ in the last bit: the split-K fc1 accumulates with atomics)."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    from probabilisticteacher_b200.modeling.postprocessing import postprocess_batch
    model = build_model(c2f_config(), cuda)
    model.init_synthetic(seed=7)
    model.eval()
    batch = O.synthetic_batch(2, 160, 224, 8, 5, labelled=False)
    out = model(batch)  # post-processed to the "height" / "width" of the inputs (= the network input size here)
    assert len(out) == 2 and out[0]["instances"].image_size == (160, 224)
    raw = model.inference(batch, do_postprocess=False)
    sizes = [(160, 224)] * 2
    p1 = postprocess_batch(raw, batch, sizes)
    p2 = postprocess_batch(raw, [dict(d, height=320, width=448) for d in batch], sizes)
    for a, b2 in zip(p1, p2):
        ia, ib = a["instances"], b2["instances"]
        assert ib.image_size == (320, 448) and ia.image_size == (160, 224)
        assert torch.equal(ib.pred_boxes.tensor, ia.pred_boxes.tensor * 2) and torch.equal(ib.scores, ia.scores)
        assert torch.equal(ib.pred_classes, ia.pred_classes) and 0 < len(ia) <= 100
