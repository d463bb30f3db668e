"""GPU parity on the other BASELINE configurations and edge cases: K = 1 (KITTI / Sim10k configs),
DefaultAnchorGenerator (base yaml), images of different sizes in one batch, eval-mode inference,
empty ground truth, and full-size (3x800x1333) kernels against torch fp32 ops / size-independent
properties."""
import pytest
import torch

pytestmark = pytest.mark.gpu


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


def _oracle_props(O, model, sizes):
    p = model._last_ctx["props"]
    out = []
    for n in range(p["boxes"].shape[0]):
        c = int(p["count"][n])
        out.append(O.OInst(sizes[n], proposal_boxes=O.OBoxes(p["boxes"][n, :c].cpu()),
                           objectness_logits=p["scores"][n, :c].cpu()))
    return out


def _check_losses(lg, lo, rpn_tol=5e-3, roi_tol=2e-2):
    for k in lo:
        tol = rpn_tol if "rpn" in k else roi_tol
        a, b = float(lg[k]), float(lo[k])
        assert abs(a - b) <= tol * max(abs(b), 1e-3), (k, a, b)


@pytest.mark.parametrize("K,anchor_gen", [(1, "DifferentiableAnchorGenerator"), (8, "DefaultAnchorGenerator")])
def test_other_configs_supervised(cuda, K, anchor_gen):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    H, W = 144, 240
    cfg = c2f_config()
    cfg.MODEL.ROI_HEADS.NUM_CLASSES = K
    cfg.MODEL.ANCHOR_GENERATOR.NAME = anchor_gen
    model = build_model(cfg, cuda)
    sd = model.init_synthetic(seed=5)
    model.train()
    om = O.OracleRCNN(O.OracleCfg(num_classes=K, anchor_generator=anchor_gen), seed=0)
    om.load_ref_state_dict(sd)
    lab = O.synthetic_batch(2, H, W, K, 11)
    g = torch.Generator().manual_seed(3)
    R = (H // 16) * (W // 16) * 9
    L = 2000 + 16
    pr = {"rpn": (torch.rand(2, R, generator=g).to(cuda), torch.rand(2, R, generator=g).to(cuda)),
          "roi": (torch.rand(2, L, generator=g).to(cuda), torch.rand(2, L, generator=g).to(cuda))}
    model.prio_override = pr
    om.sampler = _Sampler(pr)
    model.zero_grad()
    lg, _, _, _ = model(_to_inst(lab), branch="supervised")
    lo, _, _, _ = om(lab, branch="supervised", proposals_override=_oracle_props(O, model, [(H, W)] * 2))
    _check_losses(lg, lo)
    sum(lg.values()).backward()
    torch.cuda.synchronize()
    assert torch.isfinite(model.arena.grads).all()


def test_mixed_image_sizes_and_empty_gt(cuda):
    """ImageList padding (different sizes in one batch) and an image without ground truth."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    cfg = c2f_config()
    model = build_model(cfg, cuda)
    sd = model.init_synthetic(seed=6)
    model.train()
    om = O.OracleRCNN(O.OracleCfg(), seed=0)
    om.load_ref_state_dict(sd)
    a = O.synthetic_batch(1, 160, 208, 8, 21)[0]
    b = O.synthetic_batch(1, 128, 240, 8, 22)[0]
    b["instances"] = O.OInst((128, 240), gt_boxes=O.OBoxes(torch.zeros(0, 4)), gt_classes=torch.zeros(0, dtype=torch.int64))
    lab = [a, b]
    Hm, Wm = 160, 240
    g = torch.Generator().manual_seed(4)
    R = (Hm // 16) * (Wm // 16) * 9
    L = 2000 + 16
    pr = {"rpn": (torch.rand(2, R, generator=g).to(cuda), torch.rand(2, R, generator=g).to(cuda)),
          "roi": (torch.rand(2, L, generator=g).to(cuda), torch.rand(2, L, generator=g).to(cuda))}
    model.prio_override = pr
    om.sampler = _Sampler(pr)
    lg, _, _, _ = model(_to_inst(lab), branch="supervised")
    lo, _, _, _ = om(lab, branch="supervised", proposals_override=_oracle_props(O, model, [(160, 208), (128, 240)]))
    _check_losses(lg, lo)


def test_eval_mode_inference(cuda):
    from oracle import pt_oracle as O
    from probabilisticteacher_b200.config import c2f_config
    from probabilisticteacher_b200.modeling.meta_arch.rcnn import build_model
    cfg = c2f_config()
    model = build_model(cfg, cuda)
    model.init_synthetic(seed=7)
    model.eval()
    out = model(O.synthetic_batch(2, 160, 224, 8, 5, labelled=False))
    assert len(out) == 2
    inst = out[0]["instances"].trim()
    assert len(inst) <= 100 and inst.pred_boxes.tensor.shape[1] == 4
    s = inst.scores
    assert bool((s[:-1] >= s[1:]).all())  # sorted by descending score
    b = inst.pred_boxes.tensor
    assert float(b[:, 0].min()) >= 0 and float(b[:, 2].max()) <= 224 and float(b[:, 3].max()) <= 160


def test_full_size_conv_layers_vs_torch(cuda):
    """3x800x1333: conv1_2-shaped (row-window, resident filter) and block-5-shaped (generic) layers
    against torch fp32 conv2d on the same fp16-rounded operands."""
    from probabilisticteacher_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(9)
    for (H, W, cin, cout) in ((800, 1333, 64, 64), (50, 83, 512, 512)):
        x = torch.randn(1, cin, H, W, generator=g).half().to(cuda)
        w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).half().to(cuda)
        bias = torch.randn(cout, generator=g).to(cuda)
        y = ops.conv3x3(ops.to_flat(x), w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous(), bias, relu=True)
        ref = torch.relu(torch.nn.functional.conv2d(x.float(), w.float(), bias, padding=1))
        got = ops.from_flat(y).float()
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-3
        # linearity (size-independent property): conv(2x) - 2*conv(x) == -bias on the pre-activation
        y2 = ops.conv3x3(ops.to_flat(2 * x), w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous(), bias, relu=False)
        y1 = ops.conv3x3(ops.to_flat(x), w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous(), bias, relu=False)
        lin = ops.from_flat(y2).float() - 2 * ops.from_flat(y1).float() + bias.view(1, -1, 1, 1)
        assert float(lin.abs().max()) < 5e-2


def test_full_size_proposal_properties(cuda):
    """R = 37 350 anchors x 2 images: proposals sorted by score, inside the image, NMS-idempotent."""
    from oracle import pt_oracle as O
    from probabilisticteacher_b200 import ops
    from probabilisticteacher_b200.modeling.proposal_generator.proposal_utils import find_top_rpn_proposals
    g = torch.Generator().manual_seed(10)
    N, H, W, A = 2, 50, 83, 9
    R = H * W * A
    anchors = O.grid_anchors(O.differentiable_cell_anchors(torch.tensor(O.OracleCfg().anchor_wh)), H, W, 16, 0.0).to(cuda)
    lg = torch.zeros(N, H, W + 1, A)
    lg[:, :, :W] = torch.randn(N, H, W, A, generator=g)
    dl = torch.zeros(N, H, W + 1, A * 8)
    dl[:, :, :W] = torch.randn(N, H, W, A * 8, generator=g) * 0.2
    flag = torch.zeros(1, dtype=torch.int32, device=cuda)
    hw = torch.tensor([[800.0, 1333.0]] * N, device=cuda)
    boxes, scores, count = find_top_rpn_proposals(lg.reshape(N, -1, A).to(cuda), dl.reshape(N, -1, A * 8).to(cuda),
                                                  anchors, N, H, W, A, hw, 0.7, 12000, 2000, 0.0, flag)
    torch.cuda.synchronize()
    assert int(flag) == 0
    for n in range(N):
        c = int(count[n])
        assert 0 < c <= 2000
        s = scores[n, :c]
        assert bool((s[:-1] >= s[1:]).all())
        b = boxes[n, :c]
        assert float(b.min()) >= 0 and float(b[:, 2].max()) <= 1333 and float(b[:, 3].max()) <= 800
        # idempotence: NMS over the survivors keeps all of them
        order = torch.arange(c, dtype=torch.int32, device=cuda)[None]
        cap = (c + 63) // 64 * 64
        op = torch.zeros(1, cap, dtype=torch.int32, device=cuda)
        op[0, :c] = order
        ki, kc = ops.nms(b[None].contiguous(), op, torch.tensor([c], dtype=torch.int32, device=cuda), 0.7, 2000)
        assert int(kc[0]) == c
