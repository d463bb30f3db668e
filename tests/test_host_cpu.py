"""CPU: C-ABI export check, host-side containers / config / schedule, and the world_size-2 (gloo)
data-parallel plumbing."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from probabilisticteacher_b200 import _lib
    L = _lib.lib()
    protos = _lib.protos()
    assert len(protos) >= 35
    for name in protos:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(protos) <= exported


def test_no_cpu_fallback():
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    with pytest.raises(RuntimeError):
        build_model(c2f_config(), device="cpu")


def test_config_defaults_and_overrides(tmp_path):
    from probabilisticteacher_b200.config import c2f_config, get_cfg
    c = get_cfg()
    assert c.MODEL.RPN.PRE_NMS_TOPK_TRAIN == 12000 and c.UNSUPNET.TAU == [0.5, 0.5]
    base = tmp_path / "base.yaml"
    base.write_text("MODEL:\n  ROI_HEADS:\n    NUM_CLASSES: 3\nSOLVER:\n  STEPS: (100, 200)\n")
    child = tmp_path / "child.yaml"
    child.write_text('_BASE_: "base.yaml"\nUNSUPNET:\n  EMA_KEEP_RATE: 0.9996\n')
    c.merge_from_file(str(child))
    c.merge_from_list(["MODEL.ANCHOR_GENERATOR.NAME", "DifferentiableAnchorGenerator", "UNSUPNET.TAU", "[0.25,0.25]"])
    assert c.MODEL.ROI_HEADS.NUM_CLASSES == 3 and c.SOLVER.STEPS == (100, 200)
    assert c.UNSUPNET.EMA_KEEP_RATE == 0.9996 and c.UNSUPNET.TAU == [0.25, 0.25]
    assert c2f_config().MODEL.ANCHOR_GENERATOR.NAME == "DifferentiableAnchorGenerator"


def test_structures():
    from probabilisticteacher_b200.structures import Boxes, FreeInstances, Instances
    b = Boxes(torch.tensor([[0., 0., 10., 10.], [5., 5., 5., 9.]]))
    assert b.nonempty().tolist() == [True, False] and b.area().tolist() == [100.0, 0.0]
    b.clip((8, 8))
    assert b.tensor[0].tolist() == [0., 0., 8., 8.]
    i = Instances((10, 10), gt_boxes=b, gt_classes=torch.tensor([1, 2]))
    with pytest.raises(AssertionError):
        i.set("x", torch.zeros(3))
    f = FreeInstances((10, 10), gt_boxes=b)
    f.set("x", torch.zeros(3))  # no length check (pt/structures/instances.py:27-33)
    f._count = torch.tensor(1)
    assert len(f.trim().gt_boxes) == 1


def test_lr_schedule():
    """The schedule `PTrainer._optimizer_step` evaluates (solver.lr_at_iter) at the c2f config's corner points."""
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.solver import lr_at_iter
    cfg = c2f_config()
    cfg.SOLVER.BASE_LR, cfg.SOLVER.STEPS, cfg.SOLVER.GAMMA = 0.016, (30000,), 0.1
    cfg.SOLVER.WARMUP_FACTOR, cfg.SOLVER.WARMUP_ITERS, cfg.SOLVER.WARMUP_METHOD = 0.001, 400, "linear"
    cfg.SOLVER.LR_SCHEDULER_NAME = "WarmupMultiStepLR"
    assert abs(lr_at_iter(cfg, 0) - 0.016 * 0.001) < 1e-12
    assert abs(lr_at_iter(cfg, 400) - 0.016) < 1e-12
    assert abs(lr_at_iter(cfg, 30000) - 0.0016) < 1e-12


def test_state_dict_layout_conversion():
    from probabilisticteacher_b200.arena import ParamArena
    t = torch.arange(2 * 3 * 3 * 5, dtype=torch.float32).view(2, 3, 3, 5)  # [co][ky][kx][ci]
    ref = ParamArena._to_ref("conv", t, 512, 7)
    assert ref.shape == (2, 5, 3, 3) and ref[1, 4, 2, 0] == t[1, 2, 0, 4]
    f = torch.arange(2 * 49 * 4, dtype=torch.float32).view(2, 49, 4)
    r = ParamArena._to_ref("fc1", f, 4, 7)
    assert r.shape == (2, 196) and r[1, 3 * 49 + 10] == f[1, 10, 3]


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from probabilisticteacher_b200.engine import dp
rank = int(os.environ["RANK"])
dist.init_process_group("gloo", rank=rank, world_size=2)
params = torch.full((1000,), float(rank + 1))
dp.broadcast_params(params)
assert bool((params == 1.0).all())
g = torch.arange(100000, dtype=torch.float32) * (rank + 1)
works = dp.allreduce_grads(g, bucket_elems=30000, async_op=True)
for w in works:
    w.wait()
assert len(works) == 4
mean = g * dp.pre_scale()
assert torch.allclose(mean, torch.arange(100000, dtype=torch.float32) * 1.5)
# the trainer's path: ONE collective over [gradient arena | 16-float metrics tail]; the tail carries the step's 8 loss
# scalars, whose cross-rank mean is what the reference logs (pt/engine/trainer.py:394-429)
from probabilisticteacher_b200.engine.trainer import PTrainer
class _Arena: pass
class _Model: pass
tr = PTrainer.__new__(PTrainer)
tr.model = _Model(); tr.model.arena = _Arena()
a = tr.model.arena
a.grads_ext = torch.zeros(1000 + 16); a.grads = a.grads_ext[:1000]; a.metrics_tail = a.grads_ext[1000:]
a.grads.fill_(float(rank + 1))
losses = {k: torch.tensor(float(i + 1) * (rank + 1)) for i, k in enumerate(PTrainer.METRIC_KEYS)}
tr._stash_metrics(losses)
works = dp.allreduce_grads(a.grads_ext, bucket_elems=a.grads_ext.numel(), async_op=True)
assert len(works) == 1
works[0].wait()
m = tr.reduced_metrics()
assert list(m) == list(PTrainer.METRIC_KEYS)
for i, k in enumerate(PTrainer.METRIC_KEYS):
    assert abs(float(m[k]) - 1.5 * (i + 1)) < 1e-6, (k, float(m[k]))
assert bool((a.grads == 3.0).all())
# the graph step's two buckets (engine/trainer.py, _graph_body_concurrent): everything behind the backbone first, the
# backbone after the join -- together ONE all-reduce of the arena, bit for bit (SUM of two ranks is order-free)
from probabilisticteacher_b200.arena import ParamArena
pa = ParamArena(num_classes=8, differentiable_anchors=True, device="cpu", with_grads=True)
cut = pa.head_bucket_start()
gen = torch.Generator().manual_seed(100 + rank)
local = torch.randn(pa.grads.numel(), generator=gen)
one = local.clone(); dist.all_reduce(one)
two = local.clone(); dist.all_reduce(two[cut:]); dist.all_reduce(two[:cut])
assert 0 < cut < two.numel() and torch.equal(one, two)
dist.destroy_process_group()
print("ok")
'''


def test_data_parallel_plumbing_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out = p.communicate(timeout=120)[0]
        assert p.returncode == 0 and "ok" in out, out


def test_overlap_buckets_partition_the_gradient_arena():
    """The cut between the two in-graph all-reduce buckets sits exactly at the backbone boundary: every trainable
    segment before it is a VGG parameter, every segment from it on is not (RPN head, box head, predictor, anchors),
    and the cut is 128-byte aligned like every segment."""
    from probabilisticteacher_b200.arena import ParamArena
    for K, diff in ((8, True), (1, False)):
        a = ParamArena(num_classes=K, differentiable_anchors=diff, device="cpu", with_grads=True)
        cut = a.head_bucket_start()
        assert 0 < cut < a.grads.numel() and cut % 64 == 0
        n_head = 0
        for name, s in a.segments.items():
            if not s.trainable:
                continue
            o = s.offset - a.trainable_start
            if name.startswith("backbone."):
                assert o + s.numel <= cut, name
            else:
                assert o >= cut, name
                n_head += 1
        assert n_head >= 8   # rpn conv / heads, fc1, fc2, cls_score, bbox_pred (weights + biases)


def test_bucket_bounds():
    from probabilisticteacher_b200.engine.dp import bucket_bounds
    assert bucket_bounds(10, 4) == [(0, 4), (4, 8), (8, 10)]


def test_synthetic_inputs_match_the_oracle_generator():
    """bench.py's main arm draws its inputs from the package (it must not import oracle/); the generator is
    the same stream as the oracle's, so parity runs can feed both sides identical batches."""
    import torch
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.synthetic import synthetic_batch
    a = synthetic_batch(2, 64, 96, 8, 1234)
    b = O.synthetic_batch(2, 64, 96, 8, 1234)
    for da, db in zip(a, b):
        assert torch.equal(da["image"], db["image"])
        assert torch.equal(da["instances"].gt_boxes.tensor, db["instances"].gt_boxes.tensor)
        assert torch.equal(da["instances"].gt_classes, db["instances"].gt_classes)
    u = synthetic_batch(1, 64, 96, 8, 7, labelled=False)
    assert "instances" not in u[0]


def test_arena_segments_are_128_byte_aligned():
    """Every segment of the flat parameter arena starts on a 64-element boundary (128 B in the fp16 operand arena):
    TMA box rows of weight tiles must be whole L2 lines (a 48-byte misalignment of fc1 cost 1.8x on that GEMM)."""
    from probabilisticteacher_b200.arena import ParamArena
    for K, diff in ((8, True), (1, False)):
        a = ParamArena(num_classes=K, differentiable_anchors=diff, device="cpu", with_grads=False)
        for s in a.segments.values():
            assert s.offset % 64 == 0, (s.name, s.offset)
        assert a.trainable_start % 64 == 0 and a.total % 64 == 0
        # the trainable part is one contiguous suffix (all-reduce / clip / SGD run over it in one pass)
        seen_trainable = False
        for s in a.segments.values():
            if s.trainable:
                seen_trainable = True
            else:
                assert not seen_trainable, s.name


def test_k2c_config_and_lr_schedule():
    from probabilisticteacher_b200.config import c2f_config, k2c_config
    from probabilisticteacher_b200.solver import lr_at_iter
    c, k = c2f_config(), k2c_config()
    assert k.MODEL.ROI_HEADS.NUM_CLASSES == 1 and c.MODEL.ROI_HEADS.NUM_CLASSES == 8
    assert k.UNSUPNET.TAU == c.UNSUPNET.TAU == [0.5, 0.5]
    # detectron2 WarmupMultiStepLR (linear warm-up from WARMUP_FACTOR over WARMUP_ITERS, x GAMMA at each step)
    c.SOLVER.BASE_LR, c.SOLVER.STEPS, c.SOLVER.GAMMA = 0.016, (30000,), 0.1
    c.SOLVER.WARMUP_FACTOR, c.SOLVER.WARMUP_ITERS, c.SOLVER.WARMUP_METHOD = 0.001, 400, "linear"
    lr = lambda it: lr_at_iter(c, it)  # noqa: E731
    assert abs(lr(0) - 0.016 * 0.001) < 1e-12
    assert abs(lr(200) - 0.016 * (0.001 * 0.5 + 0.5)) < 1e-12
    assert lr(400) == 0.016 and lr(29999) == 0.016
    assert abs(lr(30000) - 0.0016) < 1e-12


def test_detector_postprocess():
    """d2 v0.5 `detector_postprocess`: scale to the requested output size, clip, drop empty boxes, keep every field
    aligned; fixed-capacity instances (device-side count) are trimmed first; the input is not modified."""
    from probabilisticteacher_b200.modeling.postprocessing import detector_postprocess, postprocess_batch
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    boxes = torch.tensor([[10., 20., 110., 220.], [0., 0., 400., 300.], [390., 10., 400., 10.], [5., 5., 6., 6.]])
    inst = FreeInstances((300, 400), pred_boxes=Boxes(boxes.clone()), scores=torch.tensor([.9, .8, .7, .6]),
                         pred_classes=torch.tensor([1, 2, 3, 4]), scores_logists=torch.arange(36.).view(4, 9),
                         boxes_sigma=torch.ones(4, 4))
    inst._count = torch.tensor(3)  # the 4th row is beyond the valid count
    out = detector_postprocess(inst, 600, 1200)
    assert out.image_size == (600, 1200)
    # row 2 has zero height -> dropped; row 3 was never valid
    assert out.pred_boxes.tensor.tolist() == [[30., 40., 330., 440.], [0., 0., 1200., 600.]]
    assert out.scores.tolist() == pytest.approx([.9, .8]) and out.pred_classes.tolist() == [1, 2]
    assert out.scores_logists.shape == (2, 9) and out.boxes_sigma.shape == (2, 4)
    assert torch.equal(inst.pred_boxes.tensor, boxes)
    # clipping after a down-scale with rounding overshoot
    big = FreeInstances((100, 100), pred_boxes=Boxes(torch.tensor([[-5., -5., 120., 90.]])), scores=torch.tensor([1.]))
    assert detector_postprocess(big, 50, 50).pred_boxes.tensor.tolist() == [[0., 0., 50., 45.]]
    # batch form: the output size comes from the input dict, default = network input size
    res = postprocess_batch([inst, big], [{"height": 150, "width": 200}, {}], [(300, 400), (100, 100)])
    assert res[0]["instances"].image_size == (150, 200) and res[1]["instances"].image_size == (100, 100)
    assert res[0]["instances"].pred_boxes.tensor[0].tolist() == [5., 10., 55., 110.]
    with pytest.raises(AssertionError):
        detector_postprocess(FreeInstances((10, 10), scores=torch.ones(1)), 10, 10)
    # no detection at all (device count 0 over a fixed-capacity buffer): an empty instance at the requested size
    none = FreeInstances((300, 400), pred_boxes=Boxes(torch.zeros(100, 4)), scores=torch.zeros(100),
                         pred_classes=torch.zeros(100, dtype=torch.int64))
    none._count = torch.tensor(0)
    e = detector_postprocess(none, 600, 800)
    assert len(e) == 0 and e.image_size == (600, 800) and e.pred_boxes.tensor.shape == (0, 4) and e.scores.shape == (0,)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores) at a reduced image size: one JSON
    line with the keys the driver reads; under torchrun only rank 0 prints."""
    import json
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--height", "96", "--width", "128"], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "iters/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["metric"].startswith("teacher+student training iters/sec")
    other = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                           env=dict(os.environ, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=120)
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_trainer_pseudo_label_repack_host_logic():
    """`PTrainer.threshold_bbox / process_pseudo_label / remove_label / add_label` (pt/engine/trainer.py:179-257): pure
    container shuffling, exercised without a device (the constructor, which builds the CUDA models, is skipped)."""
    from probabilisticteacher_b200.engine.trainer import PTrainer
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    tr = PTrainer.__new__(PTrainer)
    roih = FreeInstances((60, 80), pred_boxes=Boxes(torch.tensor([[1., 2., 30., 40.], [5., 5., 9., 9.]])),
                         scores=torch.tensor([.9, .2]), pred_classes=torch.tensor([3, 1]),
                         scores_logists=torch.arange(18.).view(2, 9), boxes_sigma=torch.ones(2, 4))
    roih._count = torch.tensor(2)
    out, _ = tr.process_pseudo_label([roih], "roih", "all")
    p = out[0]
    # method "all": every detection becomes a pseudo label; scores / classes are NOT carried (trainer.py:203-226)
    assert set(p.get_fields()) == {"pseudo_boxes", "scores_logists", "boxes_sigma"} and p.image_size == (60, 80)
    assert torch.equal(p.pseudo_boxes.tensor, roih.pred_boxes.tensor) and p.valid_count() is roih.valid_count()
    no_sigma = FreeInstances((60, 80), pred_boxes=roih.pred_boxes, scores_logists=roih.scores_logists)
    assert set(tr.threshold_bbox(no_sigma, "roih").get_fields()) == {"pseudo_boxes", "scores_logists"}
    rpn = FreeInstances((60, 80), proposal_boxes=Boxes(torch.tensor([[0., 0., 8., 8.]])), objectness_logits=torch.tensor([2.]))
    q = tr.threshold_bbox(rpn, "rpn")
    assert set(q.get_fields()) == {"gt_boxes", "objectness_logits", "pseudo_boxes"}
    with pytest.raises(ValueError):
        tr.process_pseudo_label([roih], "roih", "thresholding")
    data = [{"image": 0, "instances": roih}, {"image": 1}]
    assert all("instances" not in d for d in tr.remove_label(data))
    labelled = tr.add_label(data, out + out)
    assert labelled[0]["instances"] is p and labelled[1]["instances"] is p


def test_lr_schedules_closed_form():
    """`solver.lr_at_iter` against the reference's own WarmupTwoStageMultiStepLR stepping a torch optimizer
    (tests/golden/pt_reference_lr_golden.json, oracle/make_golden_lr.py), the d2 formulas of the other two
    schedulers, and the dispatch errors of pt/solver/build.py."""
    import json
    import math
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.solver import lr_at_iter
    cases = json.load(open(os.path.join(ROOT, "tests", "golden", "pt_reference_lr_golden.json")))
    for c in cases:
        cfg = c2f_config()
        cfg.SOLVER.LR_SCHEDULER_NAME = "WarmupTwoStageMultiStepLR"
        cfg.SOLVER.BASE_LR, cfg.SOLVER.STEPS, cfg.SOLVER.FACTOR_LIST = c["base_lr"], tuple(c["steps"]), tuple(c["factor_list"])
        cfg.SOLVER.WARMUP_FACTOR, cfg.SOLVER.WARMUP_ITERS, cfg.SOLVER.WARMUP_METHOD = \
            c["warmup_factor"], c["warmup_iters"], c["warmup_method"]
        for it, want in enumerate(c["lrs"]):
            assert lr_at_iter(cfg, it) == pytest.approx(want, rel=1e-12, abs=0), (c["steps"], it)
    cfg = c2f_config()
    cfg.SOLVER.LR_SCHEDULER_NAME = "WarmupCosineLR"
    assert lr_at_iter(cfg, 15000) == pytest.approx(0.016 * 0.5 * (1 + math.cos(math.pi * 0.5)), abs=1e-12)
    assert lr_at_iter(cfg, 0) == pytest.approx(0.016 * 0.001)
    cfg.SOLVER.LR_SCHEDULER_NAME = "WarmupTwoStageMultiStepLR"
    cfg.SOLVER.FACTOR_LIST = (1,)  # one milestone needs two factors
    with pytest.raises(ValueError):
        lr_at_iter(cfg, 0)
    cfg.SOLVER.LR_SCHEDULER_NAME = "Poly"
    with pytest.raises(ValueError):
        lr_at_iter(cfg, 0)


def test_train_loop_checkpoint_cadence():
    """`PTrainer.train`: fvcore PeriodicCheckpointer cadence (every CHECKPOINT_PERIOD finished iterations, named after
    the 0-based iteration that just finished, plus `model_final` at MAX_ITER) -- host logic on a stub (no device)."""
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    tr = PTrainer.__new__(PTrainer)
    cfg = c2f_config()
    cfg.SOLVER.CHECKPOINT_PERIOD, cfg.SOLVER.MAX_ITER = 4, 12
    tr.cfg, tr.iter, tr.max_iter, tr.last_losses, tr.rank = cfg, 0, 12, None, 0
    saved = []

    class _Ck:
        def save(self, name, **kw):
            saved.append((name, kw))
    tr.checkpointer = _Ck()
    checks = []

    class _Gen:
        def raise_if_nonfinite(self):
            checks.append(tr.iter)

    class _Model:
        proposal_generator = _Gen()
    tr.model, tr.model_teacher = _Model(), _Model()

    def step():
        tr.iter += 1
    tr.step = step
    tr.train(3, check_period=2)
    assert tr.iter == 3 and saved == []
    assert checks == [2, 2, 3, 3]  # both detectors, every 2nd iteration and at the end of the call
    tr.train()
    assert tr.iter == 12
    tr.last_losses = {"loss_cls": torch.tensor(float("inf"))}
    tr.iter, tr.max_iter = 0, 1
    with pytest.raises(FloatingPointError):
        tr.train(check_period=1)
    tr.max_iter = 12
    assert saved == [("model_0000003", {"iteration": 3}), ("model_0000007", {"iteration": 7}),
                     ("model_0000011", {"iteration": 11}), ("model_final", {"iteration": 11})]


def test_every_entry_point_cites_the_reference_and_is_documented():
    """include/ptb200.h is the boundary contract: every prototype sits under a comment that names the reference
    interface it replaces (a reference file with line numbers, detectron2 / torchvision behaviour, or autograd of a
    cited call), and every entry point appears in INTEGRATION.md's replacement table."""
    import re
    text = open(os.path.join(ROOT, "include", "ptb200.h")).read()
    pairs = re.findall(r"/\*((?:(?!\*/).)*)\*/\s*((?:int\s+ptb200_\w+\s*\([^;]*\)\s*;\s*)+)", text, flags=re.S)
    covered = {}
    for comment, protos in pairs:
        for name in re.findall(r"int\s+(ptb200_\w+)", protos):
            covered[name] = comment
    declared = set(re.findall(r"\bint\s+(ptb200_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    assert declared == set(covered), sorted(declared - set(covered))
    cite = re.compile(r"\.py(:\d+)?|d2 |detectron2|torchvision|torch\.sort")
    uncited = sorted(n for n, c in covered.items() if not cite.search(c))
    assert not uncited, uncited
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert not [n for n in declared if n not in doc]


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No silent fallback: without libptb200.so (or with a symbol missing) the loader raises."""
    from probabilisticteacher_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libptb200.so"))
    with pytest.raises(_lib.PTB200Error, match="no CPU fallback"):
        _lib.lib()
    # a library that lacks a declared symbol is refused as well
    src = tmp_path / "stub.c"
    src.write_text("int ptb200_ema_update(void) { return 0; }\n")
    subprocess.check_call(["gcc", "-shared", "-fPIC", "-o", str(tmp_path / "libptb200.so"), str(src)])
    with pytest.raises(_lib.PTB200Error, match="does not export"):
        _lib.lib()
    with pytest.raises(TypeError):
        monkeypatch.undo()
        _lib.call("ptb200_ema_update", 1, 2)  # wrong arity is caught before the call


def test_product_code_never_imports_the_oracle():
    """oracle/ is the checker: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it. Static scan
    of every module of the package, of tools/ and of bench.py (where the import must live inside the CPU-baseline
    function, never at module level or in the main arm)."""
    import ast
    import glob

    def oracle_imports(path):
        tree = ast.parse(open(path).read())
        hits = []
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            hits += [(m, node.lineno) for m in mods if m == "oracle" or m.startswith("oracle.")]
        return tree, hits

    files = glob.glob(os.path.join(ROOT, "probabilisticteacher_b200", "**", "*.py"), recursive=True)
    files += glob.glob(os.path.join(ROOT, "tools", "*.py"))
    assert len(files) > 20
    for f in files:
        assert oracle_imports(f)[1] == [], f
    tree, hits = oracle_imports(os.path.join(ROOT, "bench.py"))
    assert hits, "bench.py's cpu_baseline leg times the oracle"
    allowed = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "cpu_oracle_iters_per_s"]
    lo, hi = allowed[0].lineno, allowed[0].end_lineno
    assert all(lo <= line <= hi for _, line in hits), hits


def test_unsupported_freeze_at_is_refused():
    """FREEZE_AT = 0 would leave the fused, forward-only first conv in the trainable set without a gradient kernel:
    the builder refuses it instead of training silently wrong (every reference config keeps detectron2's default 2)."""
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.backbone.vgg import build_vgg_backbone
    cfg = c2f_config()
    assert cfg.MODEL.BACKBONE.FREEZE_AT == 2
    cfg.MODEL.BACKBONE.FREEZE_AT = 0
    with pytest.raises(ValueError, match="FREEZE_AT"):
        build_vgg_backbone(cfg, None, 1024.0)


def _leaves(d, prefix=""):
    for k, v in d.items():
        if isinstance(v, dict):
            yield from _leaves(v, prefix + k + ".")
        else:
            yield prefix + k, v


def _norm(v):
    import ast
    if isinstance(v, str):
        try:
            v = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            pass
    if isinstance(v, (list, tuple)):
        return [_norm(x) for x in v]
    return v


def test_config_helpers_restate_the_reference_yaml_files():
    """tests/golden/pt_reference_cfg_golden.json = the reference's own yaml files (base + child, parsed independently of
    the package's loader by oracle/make_golden_cfg.py) and train.sh's overrides. Every value they set on the hot path
    must be what `c2f_config()` / `k2c_config()` -- the configurations bench.py and the tests run -- carry."""
    import json
    from probabilisticteacher_b200.config import c2f_config, k2c_config, validate_cfg
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "pt_reference_cfg_golden.json")))
    off_path = ("DATASETS.", "DATALOADER.", "OUTPUT_DIR", "TEST.EVALUATOR", "INPUT.")  # data pipeline / bookkeeping
    for name, helper in (("final_c2f", c2f_config), ("final_k2c", k2c_config)):
        want = dict(_leaves(G[name]))
        want.update(G["train_sh"])
        cfg = validate_cfg(helper())
        have = dict(_leaves(cfg))
        checked = 0
        for k, v in want.items():
            if k.startswith(off_path):
                continue
            assert k in have, (name, k)
            assert _norm(have[k]) == _norm(v), (name, k, have[k], v)
            checked += 1
        assert checked >= 40
    # the other three reference configs differ from these two only off the hot path
    for other, same_as in (("final_c2b", "final_c2f"), ("final_c2f_0.02", "final_c2f"), ("final_s2c", "final_k2c")):
        a = {k: v for k, v in _leaves(G[other]) if not k.startswith(off_path)}
        b = {k: v for k, v in _leaves(G[same_as]) if not k.startswith(off_path)}
        assert a == b, other


@pytest.mark.skipif(not os.path.isdir("/root/reference/configs"), reason="reference tree not mounted")
def test_reference_yaml_files_load_through_the_package_config():
    """`get_cfg().merge_from_file(<the reference's own yaml>)` (with its `_BASE_` chain) + train.sh's overrides gives
    a configuration the B200 path accepts and that equals the helper on every hot-path key."""
    from probabilisticteacher_b200.config import c2f_config, get_cfg, validate_cfg
    c = get_cfg()
    c.merge_from_file("/root/reference/configs/pt/final_c2f.yaml")
    c.merge_from_list(["MODEL.ANCHOR_GENERATOR.NAME", "DifferentiableAnchorGenerator", "UNSUPNET.EFL", "True",
                       "UNSUPNET.EFL_LAMBDA", "[0.5,0.5]", "UNSUPNET.TAU", "[0.5,0.5]"])
    validate_cfg(c)
    a, b = dict(_leaves(c)), dict(_leaves(c2f_config()))
    off_path = ("DATASETS.", "DATALOADER.", "OUTPUT_DIR", "TEST.EVALUATOR")
    diff = {k for k in set(a) | set(b) if not k.startswith(off_path) and _norm(a.get(k)) != _norm(b.get(k))}
    assert not diff, diff


def test_validate_cfg_names_the_unsupported_key():
    from probabilisticteacher_b200.config import c2f_config, validate_cfg
    cases = [("MODEL.ROI_HEADS.NUM_CLASSES", 20), ("MODEL.ROI_BOX_HEAD.NUM_FC", 3), ("MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION", 14),
             ("MODEL.RPN.BBOX_REG_WEIGHTS", (1.0, 1.0, 2.0, 2.0)), ("MODEL.ROI_BOX_HEAD.CLS_AGNOSTIC_BBOX_REG", True),
             ("MODEL.BACKBONE.FREEZE_AT", 0), ("MODEL.MASK_ON", True), ("UNSUPNET.MODEL_TYPE", "LAPLACE"),
             ("MODEL.ANCHOR_GENERATOR.ANCHOR", [[[10.0, 10.0]] * 5])]
    for key, bad in cases:
        cfg = c2f_config()
        node = cfg
        *parents, leaf = key.split(".")
        for p in parents:
            node = node[p]
        node[leaf] = bad
        with pytest.raises(ValueError, match=key.split(".")[-2]):
            validate_cfg(cfg)
    cfg = c2f_config()
    cfg.MODEL.ANCHOR_GENERATOR.NAME = "DefaultAnchorGenerator"
    cfg.MODEL.ANCHOR_GENERATOR.SIZES = [[32, 64, 128, 256, 512]]
    with pytest.raises(ValueError, match="ANCHOR_GENERATOR"):
        validate_cfg(cfg)


def test_every_call_site_matches_the_header_arity():
    """Static check of the Python -> C-ABI call sites: every `call("ptb200_...", ...)` in the package passes exactly the
    parameters the header declares (the trailing `stream` may be omitted: `_lib.call` fills it in), and names an entry
    point that exists. Catches signature drift without a GPU (at run time `_lib.call` raises TypeError)."""
    import ast
    import glob
    from probabilisticteacher_b200 import _lib
    protos = _lib.protos()
    seen = set()
    for path in glob.glob(os.path.join(ROOT, "probabilisticteacher_b200", "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == "call" and node.args):
                continue
            first = node.args[0]
            if not (isinstance(first, ast.Constant) and isinstance(first.value, str) and first.value.startswith("ptb200_")):
                continue
            name = first.value
            assert name in protos, (path, node.lineno, name)
            if any(isinstance(a, ast.Starred) for a in node.args) or node.keywords:
                continue  # (none today) variadic sites cannot be counted statically
            n_args, n_decl = len(node.args) - 1, len(protos[name])
            assert n_args in (n_decl, n_decl - 1), (os.path.relpath(path, ROOT), node.lineno, name, n_args, n_decl)
            seen.add(name)
    # the tools / tests bind a few entry points through ctypes directly; most must be reachable from the package
    assert len(seen) >= 40, sorted(set(protos) - seen)
