"""CPU: the oracle's whole-model orchestration against the REFERENCE'S OWN MODEL CLASSES.

tests/golden/pt_reference_model_golden.pt was produced by oracle/make_golden_model.py: the reference's
`GuassianGeneralizedRCNN.forward` (pt/modeling/meta_arch/rcnn.py:30-92) with the reference's VGG, GuassianRPN,
(Differentiable)AnchorGenerator, GuassianROIHead and GuassianFastRCNNOutputLayers, imported unmodified and executed
end to end in the three branches of a post-burn-in iteration, for two configurations (C2F: K = 8, differentiable
anchors; K = 1 with detectron2's default anchors and two images of different sizes). Here
`oracle.pt_oracle.OracleRCNN` gets the same images, ground truth, weights (same seeded initialisers) and sampling
priorities, and must reproduce the losses, the teacher's RPN proposals and pseudo labels, parameter gradients of the
supervised branch and the gradient reaching the differentiable anchors."""
import os

import pytest
import torch

from oracle import pt_oracle as O

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_model_golden.pt"), weights_only=False)
CASES = sorted(GOLD)


class _Sampler:
    def __init__(self, pr):
        self.pr = pr

    def prio(self, tag, n):
        grp, which = tag[0].split("_")
        return self.pr[grp][0 if which == "pos" else 1][tag[1]][:n]


def _model(G):
    om = O.OracleRCNN(O.OracleCfg(num_classes=G["K"], anchor_generator=G["anchor_generator"]), seed=G["seed"])
    om.sampler = _Sampler(G["prio"])
    return om


def _batches(G):
    lab = [{"image": im, "height": hw[0], "width": hw[1],
            "instances": O.OInst(tuple(hw), gt_boxes=O.OBoxes(b.clone()), gt_classes=c.clone())}
           for im, b, c, hw in zip(G["lab_images"], G["gt_boxes"], G["gt_classes"], G["sizes"])]
    unl = [{"image": im, "height": hw[0], "width": hw[1]} for im, hw in zip(G["unl_images"], G["sizes"])]
    return lab, unl


def _close(a, b, tol=2e-5):
    a, b = torch.as_tensor(a).detach().double(), torch.as_tensor(b).detach().double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = float((a - b).abs().max()) if a.numel() else 0.0
    assert err <= tol * max(float(b.abs().max()) if b.numel() else 1.0, 1e-6), err


@pytest.mark.parametrize("case", CASES)
def test_supervised_branch_matches_reference_model(case):
    G = GOLD[case]
    om = _model(G)
    lab, _ = _batches(G)
    losses, _, _, _ = om(lab, branch="supervised")
    assert set(losses) == set(G["sup_losses"])
    for k, v in G["sup_losses"].items():
        _close(losses[k], v)
    sum(losses.values()).backward()
    _close(om.p("proposal_generator.rpn_head.objectness_logits.weight").grad, G["sup_grad_rpn_objectness_w"], 1e-4)
    _close(om.p("roi_heads.box_predictor.cls_score.weight").grad, G["sup_grad_cls_score_w"], 1e-4)
    _close(om.p("backbone.vgg_block5.0.conv3.bias").grad, G["sup_grad_conv5_3_b"], 1e-4)


@pytest.mark.parametrize("case", CASES)
def test_teacher_branch_matches_reference_model(case):
    G = GOLD[case]
    om = _model(G)
    _, unl = _batches(G)
    with torch.no_grad():
        _, props, roih, _ = om(unl, branch="unsup_data_weak")
    for n in range(G["N"]):
        _close(props[n].proposal_boxes.tensor, G["teacher_rpn_boxes"][n])
        _close(props[n].objectness_logits, G["teacher_rpn_logits"][n])
        ref = G["teacher_roih"][n]
        assert torch.equal(roih[n].pred_classes, ref["pred_classes"])
        for f in ("pred_boxes", "scores", "scores_logists", "boxes_sigma"):
            v = getattr(roih[n], f)
            _close(v.tensor if hasattr(v, "tensor") else v, ref[f])


@pytest.mark.parametrize("case", CASES)
def test_unsupervised_branch_matches_reference_model(case):
    G = GOLD[case]
    _, unl = _batches(G)
    unl_q = [dict(d, instances=O.OInst(tuple(hw), pseudo_boxes=O.OBoxes(r["pred_boxes"].clone()),
                                       scores_logists=r["scores_logists"].clone(), boxes_sigma=r["boxes_sigma"].clone()))
             for d, r, hw in zip(unl, G["teacher_roih"], G["sizes"])]
    differentiable = G["anchor_generator"] == "DifferentiableAnchorGenerator"
    for danchor, key in ((True, "unsup_anchor_grad"), (False, "unsup_anchor_grad_no_danchor")):
        om = _model(G)
        losses, _, _, _ = om(unl_q, branch="unsupervised", danchor=danchor)
        if danchor:
            assert set(losses) == set(G["unsup_losses"])
            for k, v in G["unsup_losses"].items():
                _close(losses[k], v)
        if differentiable:
            sum(losses.values()).backward()
            g = om.p("proposal_generator.anchor_generator.anchor_0").grad
            g = torch.zeros_like(G[key]) if g is None else g
            _close(g, G[key], 1e-3)
    if differentiable:  # grad_zero (pt/modeling/utils.py:47-58): anchors learn only from the unsupervised RPN branch
        assert float(G["unsup_anchor_grad"].abs().max()) > 0 and float(G["unsup_anchor_grad_no_danchor"].abs().max()) == 0


# ------------------------------------------------------------------------------------------ eval mode
EVAL = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_eval_golden.pt"), weights_only=False)


@pytest.mark.parametrize("case", sorted(EVAL))
def test_eval_mode_inference_matches_the_reference_models(case):
    """tests/golden/pt_reference_eval_golden.pt (oracle/make_golden_eval.py): the reference's own model classes in
    eval mode (test-time top-k 6000 / 1000, `GuassianFastRCNNOutputLayers.inference`, then d2's
    `detector_postprocess` to a different output size). The oracle's eval path gives the raw detections; the
    PACKAGE's `detector_postprocess` (what `GuassianGeneralizedRCNN.inference` applies on the B200 path) must turn
    them into the reference's final instances."""
    from probabilisticteacher_b200.modeling.postprocessing import detector_postprocess
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    G = EVAL[case]
    om = O.OracleRCNN(O.OracleCfg(num_classes=G["K"], anchor_generator=G["anchor_generator"]), seed=G["seed"])
    batch = [{"image": im, "height": oh, "width": ow} for im, (oh, ow) in zip(G["images"], G["out_sizes"])]
    with torch.no_grad():
        _, _, roih, _ = om(batch, branch="unsup_data_weak", training=False)
    for n, (r, ref) in enumerate(zip(roih, G["detections"])):
        _close(O._bt(r.pred_boxes), ref["raw_boxes"])
        inst = FreeInstances(tuple(G["sizes"][n]), pred_boxes=Boxes(O._bt(r.pred_boxes).clone()), scores=r.scores,
                             pred_classes=r.pred_classes, scores_logists=r.scores_logists, boxes_sigma=r.boxes_sigma)
        out = detector_postprocess(inst, *G["out_sizes"][n])
        assert tuple(out.image_size) == tuple(ref["image_size"])
        assert torch.equal(out.pred_classes, ref["pred_classes"])
        _close(out.pred_boxes.tensor, ref["pred_boxes"])
        _close(out.scores, ref["scores"])
        _close(out.scores_logists, ref["scores_logists"])
        _close(out.boxes_sigma, ref["boxes_sigma"])


# ------------------------------------------------------------------------------------------ BASELINE config 1, full size
def _config1_inputs(G):
    g = torch.Generator().manual_seed(G["prio_seed"])
    R = (G["H"] // 16) * (G["W"] // 16) * 9
    L = 2000 + 16
    N = G["N"]
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}
    lab = O.synthetic_batch(N, G["H"], G["W"], G["K"], G["lab_seed"])
    unl = O.synthetic_batch(N, G["H"], G["W"], G["K"], G["unl_seed"], labelled=False)
    return prio, lab, unl


def _overlap(a, b, scale, tol=1e-5):
    """Fraction of the rows of a that also appear in b (same box to tol * scale)."""
    d = (a.double()[:, None, :] - b.double()[None, :, :]).abs().amax(-1) / scale
    return float((d.min(1).values < tol).double().mean())


@pytest.mark.parametrize("case", ["config1", "config4"])
def test_full_size_configs_match_the_reference_models(case):
    """BASELINE.json config 1 (Guassian-RCNN-VGG.yaml, 1 source + 1 target synthetic 3x800x1333 image, one iteration's
    losses) and config 4's model and size (final_k2c.yaml: K = 1, differentiable anchors, 3x600x2000):
    tests/golden/pt_reference_{case}_golden.pt hold what the reference's own model classes compute at FULL size
    (oracle/make_golden_config1.py). (The CUDA path is compared with the same fixtures in tests/test_zz_next_rows_gpu.py.)

    config1: the oracle reproduces the 8 losses, the teacher's proposals and its 100 pseudo labels exactly.
    config4: 12 of the 12 000 best RPN logits of this input are EQUAL in fp32. The reference sorts them with
    `logits.sort(descending=True)` (proposal_utils.py:87), whose order among equal keys is unspecified; the position in
    that order picks the sigma row used for re-scoring (the :94 quirk), so the reference's own proposal list depends on
    torch's tie order. The oracle and the CUDA path break ties by ascending anchor index. Hence here: RPN losses and
    pseudo labels exact, proposal lists equal as sets up to the tied candidates (measured: 1998 of 2000), ROI losses
    (which see the sampled subset of those lists) to 1e-2."""
    strict = case == "config1"
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", f"pt_reference_{case}_golden.pt"), weights_only=False)
    prio, lab, unl = _config1_inputs(G)
    om = O.OracleRCNN(O.OracleCfg(num_classes=G["K"], anchor_generator=G["anchor_generator"]), seed=G["weight_seed"])
    om.sampler = _Sampler(prio)
    scale = float(max(G["H"], G["W"]))

    def loss_close(mine, ref, k):
        tol = 2e-5 if (strict or "rpn" in k) else 1e-2
        assert abs(float(mine) - ref) <= tol * max(abs(ref), 1e-6), (k, float(mine), ref)

    with torch.no_grad():
        ls, _, _, _ = om(lab, branch="supervised")
        for k, v in G["sup_losses"].items():
            loss_close(ls[k], v, k)
        _, props, roih, _ = om(unl, branch="unsup_data_weak")
        for n in range(G["N"]):
            if strict:
                _close(O._bt(props[n].proposal_boxes), G["teacher_rpn_boxes"][n])
                _close(props[n].objectness_logits, G["teacher_rpn_logits"][n])
            else:
                assert len(props[n].objectness_logits) == len(G["teacher_rpn_logits"][n])
                assert _overlap(O._bt(props[n].proposal_boxes), G["teacher_rpn_boxes"][n], scale) >= 0.995
            ref = G["teacher_roih"][n]
            assert torch.equal(roih[n].pred_classes, ref["pred_classes"])
            _close(O._bt(roih[n].pred_boxes), ref["pred_boxes"])
            _close(roih[n].scores, ref["scores"])
            _close(roih[n].scores_logists, ref["scores_logists"])
            _close(roih[n].boxes_sigma, ref["boxes_sigma"])
        unl_q = [dict(d, instances=O.OInst(r.image_size, pseudo_boxes=O.OBoxes(O._bt(r.pred_boxes)),
                                           scores_logists=r.scores_logists, boxes_sigma=r.boxes_sigma))
                 for d, r in zip(unl, roih)]
        lu, _, _, _ = om(unl_q, branch="unsupervised", danchor=True)
        for k, v in G["unsup_losses"].items():
            loss_close(lu[k], v, k)


@pytest.mark.parametrize("case", ["second_image_empty", "all_empty"])
def test_unsupervised_branch_without_pseudo_labels(case):
    """The "empty input" edge of the unsupervised branch: tests/golden/pt_reference_empty_pseudo_golden.pt
    (oracle/make_golden_empty_pseudo.py) holds the reference's own losses when the teacher produced no pseudo label for
    one image / for no image at all (then loss_cls = loss_box_reg = NaN, a mean over zero rois, and both RPN losses
    are 0). The oracle must agree, NaN included."""
    import math
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_empty_pseudo_golden.pt"), weights_only=False)
    g = torch.Generator().manual_seed(G["prio_seed"])
    N, R, L = G["N"], (G["H"] // 16) * (G["W"] // 16) * 9, 2000 + 16
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}
    om = O.OracleRCNN(O.OracleCfg(num_classes=G["K"]), seed=G["weight_seed"])
    om.sampler = _Sampler(prio)
    unl = O.synthetic_batch(N, G["H"], G["W"], G["K"], G["unl_seed"], labelled=False)
    c = G["cases"][case]
    q = []
    for d, r, n in zip(unl, G["teacher_roih"], c["keep"]):
        n = len(r["pred_boxes"]) if n is None else n
        q.append(dict(d, instances=O.OInst((G["H"], G["W"]), pseudo_boxes=O.OBoxes(r["pred_boxes"][:n]),
                                           scores_logists=r["scores_logists"][:n], boxes_sigma=r["boxes_sigma"][:n])))
    with torch.no_grad():
        lo, _, _, _ = om(q, branch="unsupervised", danchor=True)
    for k, v in c["losses"].items():
        if math.isnan(v):
            assert math.isnan(float(lo[k])), (k, float(lo[k]))
        else:
            assert abs(float(lo[k]) - v) <= 2e-5 * max(abs(v), 1e-6), (k, float(lo[k]), v)
    if case == "all_empty":
        assert math.isnan(c["losses"]["loss_cls"]) and c["losses"]["loss_rpn_loc"] == 0.0


def test_supervised_branch_without_any_ground_truth():
    """Same fixture: a supervised batch in which NO image has a ground-truth box (every anchor / proposal is
    background; the reference returns -0.0 for both regression losses)."""
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_empty_pseudo_golden.pt"), weights_only=False)
    g = torch.Generator().manual_seed(G["prio_seed"])
    N, R, L = G["N"], (G["H"] // 16) * (G["W"] // 16) * 9, 2000 + 16
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}
    om = O.OracleRCNN(O.OracleCfg(num_classes=G["K"]), seed=G["weight_seed"])
    om.sampler = _Sampler(prio)
    c = G["supervised_no_gt"]
    lab = O.synthetic_batch(N, G["H"], G["W"], G["K"], c["lab_seed"], boxes_per_image=0)
    with torch.no_grad():
        lo, _, _, _ = om(lab, branch="supervised")
    for k, v in c["losses"].items():
        assert abs(float(lo[k]) - v) <= 2e-5 * max(abs(v), 1e-6), (k, float(lo[k]), v)
    assert c["losses"]["loss_box_reg"] == 0.0 and c["losses"]["loss_rpn_loc"] == 0.0 and c["losses"]["loss_cls"] > 1.0


@pytest.mark.parametrize("variant", [0, 1])
def test_unsupervised_branch_other_unsupnet_settings(variant):
    """Same fixture: the unsupervised branch of the reference's own classes with EFL off / TAU [0.25, 0.25] (the value
    configs/pt/final_c2f.yaml itself carries) and with EFL_LAMBDA [1, 2] / TAU [0.5, 0.25]."""
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_empty_pseudo_golden.pt"), weights_only=False)
    v = G["unsupnet_variants"][variant]
    g = torch.Generator().manual_seed(G["prio_seed"])
    N, R, L = G["N"], (G["H"] // 16) * (G["W"] // 16) * 9, 2000 + 16
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}
    om = O.OracleRCNN(O.OracleCfg(num_classes=G["K"], efl=v["efl"], tau=tuple(v["tau"]), efl_lambda=tuple(v["efl_lambda"])),
                      seed=G["weight_seed"])
    om.sampler = _Sampler(prio)
    unl = O.synthetic_batch(N, G["H"], G["W"], G["K"], G["unl_seed"], labelled=False)
    q = [dict(d, instances=O.OInst((G["H"], G["W"]), pseudo_boxes=O.OBoxes(r["pred_boxes"]), scores_logists=r["scores_logists"],
                                   boxes_sigma=r["boxes_sigma"])) for d, r in zip(unl, G["teacher_roih"])]
    with torch.no_grad():
        lo, _, _, _ = om(q, branch="unsupervised", danchor=True)
    for k, ref in v["losses"].items():
        assert abs(float(lo[k]) - ref) <= 2e-5 * max(abs(ref), 1e-6), (k, float(lo[k]), ref)


def test_every_hyper_parameter_away_from_its_default():
    """tests/golden/pt_reference_oddcfg_golden.pt (oracle/make_golden_oddcfg.py): the reference's own model classes
    under a configuration where every honoured hyper-parameter differs from the defaults. The oracle, configured
    likewise, reproduces all 8 losses, the teacher's proposals and its (here 20) pseudo labels."""
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_oddcfg_golden.pt"), weights_only=False)
    g = torch.Generator().manual_seed(G["prio_seed"])
    N, R, L = G["N"], (G["H"] // 16) * (G["W"] // 16) * 9, G["roi_prio_len"]
    prio = {"rpn": (torch.rand(N, R, generator=g), torch.rand(N, R, generator=g)),
            "roi": (torch.rand(N, L, generator=g), torch.rand(N, L, generator=g))}
    om = O.OracleRCNN(O.OracleCfg(num_classes=G["K"], **G["oracle_kw"]), seed=G["weight_seed"])
    om.sampler = _Sampler(prio)
    lab = O.synthetic_batch(N, G["H"], G["W"], G["K"], G["lab_seed"], boxes_per_image=4)
    unl = O.synthetic_batch(N, G["H"], G["W"], G["K"], G["unl_seed"], labelled=False)
    with torch.no_grad():
        ls, _, _, _ = om(lab, branch="supervised")
        for k, v in G["sup_losses"].items():
            _close(ls[k], v)
        _, props, roih, _ = om(unl, branch="unsup_data_weak")
        for n in range(N):
            _close(O._bt(props[n].proposal_boxes), G["teacher_rpn_boxes"][n])
            ref = G["teacher_roih"][n]
            assert len(ref["scores"]) == 20 and torch.equal(roih[n].pred_classes, ref["pred_classes"])
            _close(O._bt(roih[n].pred_boxes), ref["pred_boxes"])
            _close(roih[n].scores, ref["scores"])
        q = [dict(d, instances=O.OInst(r.image_size, pseudo_boxes=O.OBoxes(O._bt(r.pred_boxes)), scores_logists=r.scores_logists,
                                       boxes_sigma=r.boxes_sigma)) for d, r in zip(unl, roih)]
        lu, _, _, _ = om(q, branch="unsupervised", danchor=True)
        for k, v in G["unsup_losses"].items():
            _close(lu[k], v)
