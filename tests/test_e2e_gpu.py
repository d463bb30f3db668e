"""GPU end-to-end parity: GuassianGeneralizedRCNN on the B200 path vs the CPU oracle on the same
synthetic batch, weights and sampling priorities (BASELINE config 1 at a reduced image size so that
the oracle finishes in seconds).

Tolerances: the B200 path computes convolutions / FC layers with fp16 operands and fp32 accumulation,
the oracle in fp32. Losses that do not depend on discrete proposal selection (RPN) must agree to 5e-3
relative; the ROI stage is compared on IDENTICAL proposals (the oracle is fed the device proposals) to
1e-2; gradients to 5e-2 of the tensor's max magnitude. Integer stages are covered bit-exactly in
tests/test_pipeline_ops_gpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

H, W, K = 192, 272, 8


def _setup(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    cfg = c2f_config()
    model = build_model(cfg, cuda)
    sd = model.init_synthetic(seed=3)
    model.train()
    om = O.OracleRCNN(O.OracleCfg(), seed=0)
    om.load_ref_state_dict(sd)
    return O, model, om, sd


def _to_inst(batch):
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    out = []
    for d in batch:
        nd = dict(d)
        if "instances" in d:
            i = d["instances"]
            nd["instances"] = FreeInstances(i.image_size, gt_boxes=Boxes(i.gt_boxes.tensor.clone()),
                                            gt_classes=i.gt_classes.clone())
        out.append(nd)
    return out


class _Sampler:
    def __init__(self, pr):
        self.pr = pr

    def prio(self, tag, n):
        grp, which = tag[0].split("_")
        return (self.pr[grp][0] if which == "pos" else self.pr[grp][1])[tag[1]].cpu()


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _oracle_props(O, model):
    p = model._last_ctx["props"]
    out = []
    for n in range(p["boxes"].shape[0]):
        c = int(p["count"][n])
        out.append(O.OInst((H, W), proposal_boxes=O.OBoxes(p["boxes"][n, :c].cpu()), objectness_logits=p["scores"][n, :c].cpu()))
    return out


def _grads(model, om):
    kinds = {s.name: s.kind for s in model.arena.segments.values()}
    og = {k.replace("__", "."): v.grad for k, v in om.named_parameters()}
    out = {}
    for name, v, gv, trainable in model.arena.exposed_parameters():
        if trainable and og.get(name) is not None:
            g = model.arena._to_ref(kinds.get(name, "mat"), gv, model.arena.C, 7).reshape(og[name].shape)
            out[name] = (_rel(g, og[name]), float(og[name].abs().max()))
    return out


def test_state_dict_round_trip(cuda):
    O, model, om, sd = _setup(cuda)
    sd2 = model.state_dict()
    assert set(sd2) == set(sd)
    for k in sd:
        assert torch.equal(sd2[k].cpu(), sd[k]), k
    assert set(k.replace("__", ".") for k in om.state_dict()) == set(sd)


def test_teacher_branch(cuda):
    O, model, om, _ = _setup(cuda)
    unl = O.synthetic_batch(2, H, W, K, 2, labelled=False)
    with torch.no_grad():
        _, pg, rg, _ = model(unl, branch="unsup_data_weak")
        props = []
        for n in range(2):
            t = pg[n].trim()
            props.append(O.OInst((H, W), proposal_boxes=O.OBoxes(t.proposal_boxes.tensor.cpu()),
                                 objectness_logits=t.objectness_logits.cpu()))
        _, po, ro, _ = om(unl, branch="unsup_data_weak", proposals_override=props)
        _, po_own, _, _ = om(unl, branch="unsup_data_weak")
    for n in range(2):
        # RPN proposals: same count within a few fp16-induced flips, best boxes agree
        assert abs(len(po_own[n].proposal_boxes) - len(props[n].proposal_boxes)) <= 0.1 * len(po_own[n].proposal_boxes)
        g = rg[n].trim()
        o = ro[n]
        assert len(g) == len(o.scores) == 100
        assert (g.scores[:20].cpu() - o.scores[:20]).abs().max() < 5e-3
        # same detections up to fp16-induced re-ordering of near-tied scores: match by class + box
        gb, gc = g.pred_boxes.tensor.cpu(), g.pred_classes.cpu()
        ob, oc = o.pred_boxes.tensor, o.pred_classes
        d = (gb[:, None, :] - ob[None, :, :]).abs().amax(-1)
        d[gc[:, None] != oc[None, :]] = 1e9
        assert (d.min(1).values < 2.0).float().mean() > 0.8


def test_supervised_branch_losses_and_grads(cuda):
    O, model, om, _ = _setup(cuda)
    lab = O.synthetic_batch(2, H, W, K, 1)
    g = torch.Generator().manual_seed(7)
    R = (H // 16) * (W // 16) * 9
    L = 2000 + 16
    pr = {"rpn": (torch.rand(2, R, generator=g).to(cuda), torch.rand(2, R, generator=g).to(cuda)),
          "roi": (torch.rand(2, L, generator=g).to(cuda), torch.rand(2, L, generator=g).to(cuda))}
    model.prio_override = pr
    om.sampler = _Sampler(pr)
    model.zero_grad()
    lg, _, _, _ = model(_to_inst(lab), branch="supervised")
    lo, _, _, _ = om(lab, branch="supervised", proposals_override=_oracle_props(O, model))
    for k in ("loss_rpn_cls", "loss_rpn_loc"):
        assert abs(float(lg[k]) - float(lo[k])) <= 5e-3 * abs(float(lo[k])), (k, float(lg[k]), float(lo[k]))
    for k in ("loss_cls", "loss_box_reg"):
        assert abs(float(lg[k]) - float(lo[k])) <= 1e-2 * abs(float(lo[k])), (k, float(lg[k]), float(lo[k]))
    sum(lg.values()).backward()
    sum(lo.values()).backward()
    torch.cuda.synchronize()
    for name, (r, m) in _grads(model, om).items():
        assert r < 5e-2, (name, r, m)


def test_unsupervised_branch_losses_and_grads(cuda):
    O, model, om, _ = _setup(cuda)
    unl = O.synthetic_batch(2, H, W, K, 2, labelled=False)
    from probabilisticteacher_b200.structures import Boxes, FreeInstances
    with torch.no_grad():
        _, _, roih, _ = om(unl, branch="unsup_data_weak")
    pseudo = [O.OInst(r.image_size, pseudo_boxes=O.OBoxes(r.pred_boxes.tensor), scores_logists=r.scores_logists,
                      boxes_sigma=r.boxes_sigma) for r in roih]
    unl_o = [dict(d, instances=p) for d, p in zip(unl, pseudo)]
    unl_g = [dict(d, instances=FreeInstances(p.image_size, pseudo_boxes=Boxes(p.pseudo_boxes.tensor.to(cuda)),
                                             scores_logists=p.scores_logists.to(cuda), boxes_sigma=p.boxes_sigma.to(cuda)))
             for d, p in zip(unl, pseudo)]
    model.zero_grad()
    lg, _, _, _ = model(unl_g, branch="unsupervised", danchor=True)
    lo, _, _, _ = om(unl_o, branch="unsupervised", danchor=True, proposals_override=_oracle_props(O, model))
    for k in ("loss_rpn_cls", "loss_rpn_loc"):
        assert abs(float(lg[k]) - float(lo[k])) <= 5e-3 * abs(float(lo[k])), (k, float(lg[k]), float(lo[k]))
    for k in ("loss_cls", "loss_box_reg"):
        assert abs(float(lg[k]) - float(lo[k])) <= 1e-2 * abs(float(lo[k])), (k, float(lg[k]), float(lo[k]))
    sum(lg.values()).backward()
    sum(lo.values()).backward()
    torch.cuda.synchronize()
    gr = _grads(model, om)
    for name, (r, m) in gr.items():
        assert r < 5e-2, (name, r, m)
    assert gr["proposal_generator.anchor_generator.anchor_0"][1] > 0  # danchor=True reaches the anchor parameter


def test_trainer_steps_and_ema(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 1
    lab = O.synthetic_batch(2, H, W, K, 1)
    unl = O.synthetic_batch(2, H, W, K, 2, labelled=False)

    def loader():
        while True:
            yield _to_inst(lab), _to_inst(lab), _to_inst(unl), _to_inst(unl)
    tr = PTrainer(cfg, loader(), device=cuda, seed=3)
    p0 = tr.model.arena.data.clone()
    l0 = tr.run_step()  # burn-in step (supervised only)
    assert set(l0) == {"loss_cls", "loss_box_reg", "loss_rpn_cls", "loss_rpn_loc"}
    assert not torch.equal(tr.model.arena.data, p0)
    frozen = tr.model.arena.trainable_start
    assert torch.equal(tr.model.arena.data[:frozen], p0[:frozen])  # FREEZE_AT = 2
    l1 = tr.run_step()  # iter == BURN_UP_STEP: teacher <- student copy, full PT iteration
    assert len(l1) == 8 and all(torch.isfinite(v).all() for v in l1.values())
    t_before = tr.model_teacher.arena.data.clone()
    s_before = tr.model.arena.data.clone()
    tr.run_step()       # EMA with keep rate 0.9996 happens at the start of this step
    expect = s_before * (1 - 0.9996) + t_before * 0.9996
    assert torch.allclose(tr.model_teacher.arena.data, expect, atol=1e-6)


@pytest.mark.parametrize("concurrent", [False, True])
def test_cuda_graph_step_matches_eager(cuda, concurrent):
    """The captured step (PTrainer.run_step_graphed) must reproduce the eager step: same losses (up to
    fp32 atomic-accumulation order) with identical resize geometry and sampling priorities."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    lab = O.synthetic_batch(2, H, W, K, 1)
    unl = O.synthetic_batch(2, H, W, K, 2, labelled=False)

    def loader():
        while True:
            yield _to_inst(lab), _to_inst(lab), _to_inst(unl), _to_inst(unl)
    g = torch.Generator().manual_seed(7)
    R = (H // 16) * (W // 16) * 9
    trainers = []
    for use_graph in (False, True):
        tr = PTrainer(cfg, loader(), device=cuda, seed=3, use_cuda_graph=use_graph, graph_warmup=1, gt_capacity=16,
                      concurrent=concurrent and use_graph)
        trainers.append(tr)
    L = 2000 + 16
    pr = {"rpn": (torch.rand(4, R, generator=g).to(cuda), torch.rand(4, R, generator=g).to(cuda)),
          "roi": (torch.rand(4, L, generator=g).to(cuda), torch.rand(4, L, generator=g).to(cuda))}
    for tr in trainers:
        tr.model.prio_override = pr
    hist = [[], []]
    for step in range(4):
        for i, tr in enumerate(trainers):
            losses = tr.step()
            torch.cuda.synchronize()
            hist[i].append({k: float(v) for k, v in losses.items()})
    assert trainers[1]._graph is not None
    print(hist)
    # fp32 atomic accumulation order differs run to run, and proposal selection / roi sampling are
    # discrete, so trajectories drift apart slowly: RPN losses (no discrete dependence on the other
    # branch) must agree tightly, ROI losses loosely, over the first graphed steps.
    # step 0: both eager (identical up to atomics); step 1: eager vs the graph body run eagerly on the
    # static buffers; step 2: eager vs the first captured replay; step 3: second replay (loose: chaos)
    tols = [(1e-3, 1e-3), (2e-2, 0.1), (0.1, 0.3), (1.0, 1.0)]
    for (t_rpn, t_roi), a, b in zip(tols, hist[0], hist[1]):
        for k in a:
            tol = t_rpn if "rpn" in k else t_roi
            assert abs(a[k] - b[k]) <= tol * max(abs(a[k]), 1e-3), (k, a[k], b[k])
            assert b[k] == b[k]


def test_graph_step_without_host_sync_sees_each_steps_own_ground_truth(cuda):
    """The host runs ahead of the device in graph mode (a replay is launched in microseconds): the pinned staging of
    ground truth / resize geometry must not be rewritten while an earlier step's host->device copy is still queued.
    With lr = 0 the weights never change, so every step's supervised losses are a function of THAT step's inputs only:
    the graphed trainer, driven WITHOUT any host synchronisation and with different ground truth every step, must
    reproduce the eager trainer's per-step losses."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.SOLVER.BASE_LR = 0.0
    cfg.SOLVER.WEIGHT_DECAY = 0.0
    unl = O.synthetic_batch(2, H, W, K, 2, labelled=False)
    labs = [O.synthetic_batch(2, H, W, K, 100 + s, boxes_per_image=1 + (3 * s) % 7) for s in range(8)]

    def loader():
        s = 0
        while True:
            lab = labs[s % len(labs)]
            yield _to_inst(lab), _to_inst(lab), _to_inst(unl), _to_inst(unl)
            s += 1
    g = torch.Generator().manual_seed(7)
    R = (H // 16) * (W // 16) * 9
    pr = {"rpn": (torch.rand(4, R, generator=g).to(cuda), torch.rand(4, R, generator=g).to(cuda)),
          "roi": (torch.rand(4, 2016, generator=g).to(cuda), torch.rand(4, 2016, generator=g).to(cuda))}
    hist = []
    for use_graph in (False, True):
        tr = PTrainer(cfg, loader(), device=cuda, seed=3, use_cuda_graph=use_graph, graph_warmup=1, gt_capacity=16)
        tr.model.prio_override = pr
        rec = []
        for step in range(8):
            losses = tr.step()
            rec.append({k: v.clone() for k, v in losses.items() if k.endswith("_sup")})  # stream-ordered copy, no sync
        torch.cuda.synchronize()
        assert (tr._graph is not None) == use_graph
        hist.append([{k: float(v) for k, v in r.items()} for r in rec])
    for s, (a, b) in enumerate(zip(*hist)):
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-4 * max(abs(a[k]), 1e-3), (s, k, a[k], b[k])
    # the inputs really differ from step to step
    assert len({round(h["loss_rpn_loc_sup"], 5) for h in hist[0]}) >= 6


@pytest.mark.parametrize("concurrent", [False, True])
def test_graph_step_honours_teacher_update_iter(cuda, concurrent):
    """pt/engine/trainer.py:296-298: the teacher is refreshed only when (iter - BURN_UP_STEP) % TEACHER_UPDATE_ITER
    == 0. The captured step reads the keep rate from device memory (1.0 on the other iterations: bit-exact no-op)."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.engine.trainer import PTrainer
    cfg = c2f_config()
    cfg.UNSUPNET.BURN_UP_STEP = 0
    cfg.UNSUPNET.TEACHER_UPDATE_ITER = 2
    cfg.UNSUPNET.EMA_KEEP_RATE = 0.9
    lab = O.synthetic_batch(2, H, W, K, 1)
    unl = O.synthetic_batch(2, H, W, K, 2, labelled=False)

    def loader():
        while True:
            yield _to_inst(lab), _to_inst(lab), _to_inst(unl), _to_inst(unl)
    tr = PTrainer(cfg, loader(), device=cuda, seed=3, use_cuda_graph=True, graph_warmup=1, gt_capacity=16,
                  concurrent=concurrent)
    for it in range(7):
        t_before = tr.model_teacher.arena.data.clone()
        s_before = tr.model.arena.data.clone()
        tr.step()
        torch.cuda.synchronize()
        t_after = tr.model_teacher.arena.data
        if it == 0:
            assert torch.equal(t_after, s_before)                      # iter == BURN_UP_STEP: copy (eager step)
        elif it % 2 == 0:
            expect = 0.9 * t_before.double() + (1 - 0.9) * s_before.double()
            # fp32 rounding of the two products (anchor parameters are ~500; the terms may cancel)
            bound = 1e-6 * (0.9 * t_before.double().abs() + 0.1 * s_before.double().abs()) + 1e-12
            worst = float(((t_after.double() - expect).abs() / bound).max())
            assert worst <= 1.0, (it, worst)
            assert not torch.equal(t_after, t_before)
        else:
            assert torch.equal(t_after, t_before), it                   # skipped: untouched bit for bit
    assert tr._graph is not None
