"""CPU: checkpoint I/O in the reference's on-disk format (`pt/checkpoint/detection_checkpoint.py`,
`pt/modeling/meta_arch/ts_ensemble.py`, the `vgg16_caffe.pth` key map of `pt/modeling/backbone/vgg.py:127-152`).
The arena's device passes (`pack`) are CUDA-only, so the host logic is exercised on a CPU arena whose `pack` is a
no-op; the same round trip runs on the real model in tests/test_zz_next_rows_gpu.py."""
import os

import pytest
import torch

from probabilisticteacher_b200 import checkpoint as C
from probabilisticteacher_b200.arena import ParamArena


class _CpuDetector:
    """state_dict / load_state_dict of GuassianGeneralizedRCNN (modeling/meta_arch/rcnn.py) over a CPU arena."""

    def __init__(self, K=8, diff=True, with_grads=True, seed=None):
        self.arena = ParamArena(num_classes=K, differentiable_anchors=diff, device="cpu", with_grads=with_grads)
        self.arena.pack = lambda dgrad=None: None
        if seed is not None:
            g = torch.Generator().manual_seed(seed)
            self.arena.data.copy_(torch.randn(self.arena.total, generator=g))
            # padding between / inside segments is not part of any parameter: keep it zero as the real arena does
            sd = self.arena.state_dict()
            self.arena.data.zero_()
            self.arena.load_state_dict(sd)

    def state_dict(self):
        return self.arena.state_dict()

    def load_state_dict(self, sd, strict=True):
        from torch.nn.modules.module import _IncompatibleKeys
        return _IncompatibleKeys(*self.arena.load_state_dict(sd, strict=strict))


class _Trainer:
    def __init__(self, student, it):
        self.model = student
        self.iter = it


def test_vgg16_caffe_key_map_is_the_reference_table():
    m = C.vgg16_caffe_key_map(prefix="")
    # the two literal lists of vgg.py:129-147, spot-checked at both ends and at every block boundary
    assert len(m) == 26
    assert m["features.0.weight"] == "vgg_block1.0.conv1.weight" and m["features.2.bias"] == "vgg_block1.0.conv2.bias"
    assert m["features.5.weight"] == "vgg_block2.0.conv1.weight" and m["features.7.weight"] == "vgg_block2.0.conv2.weight"
    assert m["features.10.weight"] == "vgg_block3.0.conv1.weight" and m["features.14.bias"] == "vgg_block3.0.conv3.bias"
    assert m["features.17.weight"] == "vgg_block4.0.conv1.weight" and m["features.21.weight"] == "vgg_block4.0.conv3.weight"
    assert m["features.24.weight"] == "vgg_block5.0.conv1.weight" and m["features.28.bias"] == "vgg_block5.0.conv3.bias"
    assert list(m)[:4] == ["features.0.weight", "features.0.bias", "features.2.weight", "features.2.bias"]


def test_state_dict_has_the_reference_names_and_shapes():
    sd = _CpuDetector(K=8, diff=True, with_grads=False).state_dict()
    want = {
        "backbone.vgg_block1.0.conv1.weight": (64, 3, 3, 3), "backbone.vgg_block5.0.conv3.bias": (512,),
        "proposal_generator.rpn_head.conv.weight": (512, 512, 3, 3),
        "proposal_generator.rpn_head.objectness_logits.weight": (9, 512, 1, 1),
        "proposal_generator.rpn_head.anchor_deltas.weight": (72, 512, 1, 1),
        "proposal_generator.rpn_head.anchor_deltas.bias": (72,),
        "proposal_generator.anchor_generator.anchor_0": (9, 2),
        "roi_heads.box_head.fc1.weight": (1024, 25088), "roi_heads.box_head.fc2.weight": (1024, 1024),
        "roi_heads.box_predictor.cls_score.weight": (9, 1024), "roi_heads.box_predictor.bbox_pred.weight": (64, 1024),
        "roi_heads.box_predictor.bbox_pred.bias": (64,),
    }
    for k, shp in want.items():
        assert tuple(sd[k].shape) == shp, (k, tuple(sd[k].shape))
    assert len(sd) == 26 + 6 + 1 + 8  # 13 convs, RPN head (3 layers), anchors, fc1/fc2/cls_score/bbox_pred
    assert "proposal_generator.anchor_generator.anchor_0" not in _CpuDetector(K=1, diff=False, with_grads=False).state_dict()


def test_ensemble_round_trip_through_a_file(tmp_path):
    teacher, student = _CpuDetector(with_grads=False, seed=1), _CpuDetector(seed=2)
    student.arena.momentum.copy_(torch.randn(student.arena.momentum.numel(), generator=torch.Generator().manual_seed(3)))
    # momentum padding is never touched by the optimizer either
    msd = student.arena.momentum_state_dict()
    student.arena.load_momentum_state_dict(msd)
    ck = C.DetectionTSCheckpointer(C.EnsembleTSModel(teacher, student), str(tmp_path),
                                   optimizer=C.ArenaSGDState(_Trainer(student, 41)))
    path = ck.save("model_0000040", iteration=40)
    assert os.path.basename(path) == "model_0000040.pth" and ck.has_checkpoint()
    assert ck.get_checkpoint_file() == path
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert set(raw) == {"model", "optimizer", "iteration"}
    keys = list(raw["model"])
    assert all(k.startswith("modelTeacher.") or k.startswith("modelStudent.") for k in keys)
    assert "modelStudent.roi_heads.box_head.fc1.weight" in keys and len(keys) == 2 * 41

    t2, s2 = _CpuDetector(with_grads=False, seed=7), _CpuDetector(seed=8)
    ck2 = C.DetectionTSCheckpointer(C.EnsembleTSModel(t2, s2), str(tmp_path), optimizer=C.ArenaSGDState(_Trainer(s2, 0)))
    rest = ck2.resume_or_load("", resume=True)
    assert rest["iteration"] == 40
    assert torch.equal(t2.arena.data, teacher.arena.data) and torch.equal(s2.arena.data, student.arena.data)
    assert torch.equal(s2.arena.momentum, student.arena.momentum)
    inc = ck2.last_incompatible
    assert inc.missing_keys == [] and inc.unexpected_keys == [] and inc.incorrect_shapes == []


def test_weights_only_load_ignores_optimizer_state(tmp_path):
    teacher, student = _CpuDetector(with_grads=False, seed=1), _CpuDetector(seed=2)
    student.arena.momentum.fill_(1.0)
    ck = C.DetectionTSCheckpointer(C.EnsembleTSModel(teacher, student), str(tmp_path / "a"),
                                   optimizer=C.ArenaSGDState(_Trainer(student, 5)))
    path = ck.save("model_final", iteration=4)
    t2, s2 = _CpuDetector(with_grads=False), _CpuDetector()
    ck2 = C.DetectionTSCheckpointer(C.EnsembleTSModel(t2, s2), str(tmp_path / "b"), optimizer=C.ArenaSGDState(_Trainer(s2, 0)))
    # resume=False (or nothing to resume from): only the weights of MODEL.WEIGHTS (trainer.py:483-486)
    rest = ck2.resume_or_load(path, resume=True)
    assert torch.equal(s2.arena.data, student.arena.data) and float(s2.arena.momentum.abs().sum()) == 0.0
    assert "optimizer" in rest  # left unconsumed, as fvcore returns it
    assert ck2.load("") == {}
    with pytest.raises(AssertionError):
        ck2.load(str(tmp_path / "missing.pth"))


def test_shape_mismatch_is_dropped_and_reported_and_module_prefix_is_stripped(tmp_path):
    src = _CpuDetector(K=8, seed=4)
    sd = {"module." + k: v for k, v in src.state_dict().items()}  # saved from a DDP-wrapped model
    torch.save({"model": sd}, tmp_path / "k8.pth")
    dst = _CpuDetector(K=1, diff=True, seed=5)
    before = dst.arena.state_dict()
    ck = C.DetectionTSCheckpointer(dst, str(tmp_path))
    ck.load(str(tmp_path / "k8.pth"))
    inc = ck.last_incompatible
    bad = {k for k, _, _ in inc.incorrect_shapes}
    assert bad == {"roi_heads.box_predictor.cls_score.weight", "roi_heads.box_predictor.cls_score.bias",
                   "roi_heads.box_predictor.bbox_pred.weight", "roi_heads.box_predictor.bbox_pred.bias"}
    assert set(inc.missing_keys) == bad and inc.unexpected_keys == []
    after = dst.arena.state_dict()
    for k in bad:
        assert torch.equal(after[k], before[k])  # untouched
    assert torch.equal(after["roi_heads.box_head.fc1.weight"], src.state_dict()["roi_heads.box_head.fc1.weight"])


def test_caffe2_authored_file_updates_the_student_only(tmp_path):
    teacher, student = _CpuDetector(with_grads=False, seed=1), _CpuDetector(seed=2)
    t_before = teacher.arena.data.clone()
    donor = _CpuDetector(seed=9)
    torch.save({"model": donor.state_dict(), "__author__": "Caffe2"}, tmp_path / "pre.pth")
    ck = C.DetectionTSCheckpointer(C.EnsembleTSModel(teacher, student), str(tmp_path))
    ck.load(str(tmp_path / "pre.pth"), checkpointables=[])
    assert torch.equal(student.arena.data, donor.arena.data) and torch.equal(teacher.arena.data, t_before)


def test_load_vgg16_caffe(tmp_path):
    g = torch.Generator().manual_seed(0)
    chans = [3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512]
    tv = {}
    for i, idx in enumerate(C._VGG16_FEATURE_IDX):
        tv[f"features.{idx}.weight"] = torch.randn(chans[i + 1], chans[i], 3, 3, generator=g)
        tv[f"features.{idx}.bias"] = torch.randn(chans[i + 1], generator=g)
    tv["classifier.0.weight"] = torch.zeros(8, 8)  # ignored, as in vgg.py:148-151
    torch.save(tv, tmp_path / "vgg16_caffe.pth")
    m = _CpuDetector(seed=3)
    fc1 = m.state_dict()["roi_heads.box_head.fc1.weight"].clone()
    inc = C.load_vgg16_caffe(m, str(tmp_path / "vgg16_caffe.pth"))
    sd = m.state_dict()
    assert torch.equal(sd["backbone.vgg_block1.0.conv1.weight"], tv["features.0.weight"])
    assert torch.equal(sd["backbone.vgg_block3.0.conv3.weight"], tv["features.14.weight"])
    assert torch.equal(sd["backbone.vgg_block5.0.conv3.bias"], tv["features.28.bias"])
    assert torch.equal(sd["roi_heads.box_head.fc1.weight"], fc1)
    assert inc.incorrect_shapes == [] and "roi_heads.box_head.fc1.weight" in inc.missing_keys
    with pytest.raises(KeyError):
        C.load_vgg16_caffe(m, {"features.0.weight": tv["features.0.weight"]})


def test_momentum_layout_conversion_matches_the_parameter_conversion():
    """A momentum buffer must travel through the same layout conversion as its parameter (conv [Co][ky][kx][Ci]
    <-> [Co][Ci][3][3], fc1 [1024][49][512] <-> [1024][512*49], fused head blocks split into the reference's layers)."""
    m = _CpuDetector(seed=6)
    a = m.arena
    # make the momentum arena a copy of the trainable parameters: both state dicts must then agree key by key
    a.momentum.copy_(a.data[a.trainable_start:])
    sd, msd = a.state_dict(), a.momentum_state_dict()
    assert "backbone.vgg_block1.0.conv1.weight" not in msd and "backbone.vgg_block2.0.conv2.bias" not in msd  # frozen
    assert len(msd) == 41 - 8
    for k, v in msd.items():
        assert torch.equal(v, sd[k]), k
    a.momentum.zero_()
    a.load_momentum_state_dict(msd)
    ref = a.data[a.trainable_start:].clone()
    assert torch.equal(a.momentum, ref)


def test_strip_prefix_only_when_every_key_has_it():
    sd = {"module.a": 1, "b": 2}
    assert C.strip_prefix_if_present(dict(sd), "module.") == sd
    assert C.strip_prefix_if_present({"module.a": 1, "module.b": 2}, "module.") == {"a": 1, "b": 2}


def test_key_names_are_those_of_the_reference_models_own_state_dict():
    """tests/golden/pt_reference_step_golden.pt stores samples of every parameter of the reference's OWN
    GuassianGeneralizedRCNN (built by oracle/make_golden_step.py from the unmodified pt/modeling classes) under its
    `named_parameters` keys: a checkpoint written here must use exactly those names."""
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_step_golden.pt"), weights_only=False)
    ref_names = set(G["steps"][0]["student"])
    assert set(_CpuDetector(K=G["K"], diff=True, with_grads=False).state_dict()) == ref_names
    ens = C.EnsembleTSModel(_CpuDetector(K=G["K"], with_grads=False), _CpuDetector(K=G["K"], with_grads=False))
    assert set(ens.state_dict()) == {p + k for p in ("modelTeacher.", "modelStudent.") for k in ref_names}


def test_vgg16_caffe_key_map_against_the_reference_constructor():
    """tests/golden/pt_reference_burnin_golden.pt records where the reference's OWN `VGG.__init__`
    (pt/modeling/backbone/vgg.py:127-152) put each `features.N.*` tensor of a torchvision-style file."""
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_burnin_golden.pt"), weights_only=False)
    assert dict(C.vgg16_caffe_key_map(prefix="")) == G["vgg16_caffe_mapping"]


def test_untrusted_pickles_are_refused(tmp_path):
    """Checkpoints are read with weights_only=True; a file that needs full unpickling is refused unless the caller
    says it trusts it."""
    import numpy as np
    m = _CpuDetector(seed=1)
    sd = {k: (v.numpy() if k.endswith("fc2.bias") else v) for k, v in m.state_dict().items()}  # a numpy entry
    torch.save({"model": sd}, tmp_path / "zoo.pth")
    ck = C.DetectionTSCheckpointer(_CpuDetector(seed=2), str(tmp_path))
    with pytest.raises(RuntimeError, match="trusted=True"):
        ck.load(str(tmp_path / "zoo.pth"))
    ck.load(str(tmp_path / "zoo.pth"), trusted=True)  # numpy arrays become tensors (_convert_ndarray_to_tensor)
    assert torch.equal(ck.model.state_dict()["roi_heads.box_head.fc2.bias"], m.state_dict()["roi_heads.box_head.fc2.bias"])
    assert isinstance(np.zeros(1), np.ndarray)


def test_optimizer_entry_is_interchangeable_with_torch_sgd(tmp_path):
    """The `optimizer` entry travels both ways (trainer.py:104-111 saves torch SGD's state_dict): what this trainer
    writes loads into a torch.optim.SGD built over the reference's parameter order, and a state_dict written by
    torch SGD (integer keys, no names) loads into the momentum arena by position, shapes checked."""
    student = _CpuDetector(seed=2)
    student.arena.momentum.copy_(torch.randn(student.arena.momentum.numel(), generator=torch.Generator().manual_seed(3)))
    student.arena.load_momentum_state_dict(student.arena.momentum_state_dict())  # zero the padding
    st = C.ArenaSGDState(_Trainer(student, 7)).state_dict()
    order = C.reference_trainable_order(student.arena)
    assert len(order) == 33 and order[20:24] == ["proposal_generator.rpn_head.objectness_logits.weight",
                                                  "proposal_generator.rpn_head.objectness_logits.bias",
                                                  "proposal_generator.rpn_head.anchor_deltas.weight",
                                                  "proposal_generator.rpn_head.anchor_deltas.bias"]
    sd = student.state_dict()
    params = [torch.nn.Parameter(sd[n].clone()) for n in order]
    opt = torch.optim.SGD([{"params": [p]} for p in params], lr=0.01, momentum=0.9, weight_decay=1e-4)
    torch.save({"optimizer": st}, tmp_path / "o.pth")
    opt.load_state_dict(torch.load(tmp_path / "o.pth", weights_only=True)["optimizer"])  # torch accepts the entry
    msd = student.arena.momentum_state_dict()
    for p, n in zip(params, order):
        assert torch.equal(opt.state[p]["momentum_buffer"], msd[n]), n
    # ... and back: a file written by torch SGD alone
    ref_written = opt.state_dict()
    assert set(ref_written) == {"state", "param_groups"}
    other = _CpuDetector(seed=9)
    C.ArenaSGDState(_Trainer(other, 0)).load_state_dict(ref_written)
    assert torch.equal(other.arena.momentum, student.arena.momentum)
    # a mismatching parameter count is refused instead of silently zeroing the momentum
    bad = {"state": {}, "param_groups": ref_written["param_groups"][:-1]}
    with pytest.raises(ValueError):
        C.ArenaSGDState(_Trainer(other, 0)).load_state_dict(bad)
    with pytest.raises(ValueError):
        C.ArenaSGDState(_Trainer(other, 0)).load_state_dict({"foo": 1})
