"""CPU: analytic / third-party pins of the detectron2-level behaviour the oracle restates
(the reference ships no tests, these are the pins created for it)."""
import math

import torch

from oracle import pt_oracle as O


def test_nms_matches_torchvision():
    from torchvision.ops import nms as tv_nms
    g = torch.Generator().manual_seed(0)
    for n in (1, 17, 500, 3000):
        xy = torch.rand(n, 2, generator=g) * 300
        wh = torch.rand(n, 2, generator=g) * 120 + 4
        b = torch.cat([xy, xy + wh], 1)
        s = torch.rand(n, generator=g)
        for thr in (0.5, 0.7):
            assert torch.equal(O.nms(b, s, thr), tv_nms(b, s, thr))


def test_batched_nms_matches_torchvision_vanilla():
    from torchvision.ops import boxes as tvb
    g = torch.Generator().manual_seed(1)
    n = 800
    xy = torch.rand(n, 2, generator=g) * 300
    wh = torch.rand(n, 2, generator=g) * 120 + 4
    b = torch.cat([xy, xy + wh], 1)
    s = torch.rand(n, generator=g)
    idx = torch.randint(0, 8, (n,), generator=g)
    ref = tvb._batched_nms_vanilla(b, s, idx, 0.5)
    assert torch.equal(O.batched_nms(b, s, idx, 0.5), ref)


def test_default_anchor_known_answer():
    # detectron2 tests/modeling/test_anchor_generator.py known answer (sizes 32,64; ratios .25,1,4; stride 4)
    cell = O.default_cell_anchors((32, 64), (0.25, 1, 4))
    a = O.grid_anchors(cell, 1, 2, 4, 0.0)
    expect = torch.tensor([[-32., -8., 32., 8.], [-16., -16., 16., 16.], [-8., -32., 8., 32.],
                           [-64., -16., 64., 16.], [-32., -32., 32., 32.], [-16., -64., 16., 64.],
                           [-28., -8., 36., 8.], [-12., -16., 20., 16.], [-4., -32., 12., 32.],
                           [-60., -16., 68., 16.], [-28., -32., 36., 32.], [-12., -64., 20., 64.]])
    assert torch.allclose(a, expect)


def test_box_transform_round_trip():
    a = torch.tensor([[0., 0., 16., 16.], [8., 8., 40., 24.]])
    b = torch.tensor([[2., 1., 20., 18.], [0., 0., 50., 30.]])
    d = O.get_deltas(a, b, (1., 1., 1., 1.))
    assert torch.allclose(d, torch.tensor([[0.1875, 0.0938, 0.1178, 0.0606], [0.0312, -0.0625, 0.4463, 0.6286]]), atol=1e-4)
    assert torch.allclose(O.apply_deltas(d, a, (1., 1., 1., 1.)), b, atol=1e-4)


def test_matcher_semantics():
    iou = torch.tensor([[0.1, 0.5, 0.8, 0.0], [0.2, 0.2, 0.75, 0.0]])
    m, lab = O.matcher(iou, (0.3, 0.7), (0, -1, 1), True)
    assert m.tolist() == [1, 0, 0, 0]
    assert lab.tolist()[:3] == [0, -1, 1]
    m, lab = O.matcher(torch.zeros(0, 5), (0.3, 0.7), (0, -1, 1), True)
    assert m.tolist() == [0] * 5 and lab.tolist() == [0] * 5
    # a gt whose best IoU is 0 promotes every zero-IoU prediction (detectron2 low-quality quirk)
    m, lab = O.matcher(torch.tensor([[0.0, 0.0, 0.0]]), (0.3, 0.7), (0, -1, 1), True)
    assert lab.tolist() == [1, 1, 1]


def test_subsample_labels_counts():
    g = torch.Generator().manual_seed(3)
    lab = torch.randint(-1, 2, (1000,), generator=g)
    pos, neg = O.subsample_labels(lab, 256, 0.25, 0, torch.rand(1000, generator=g), torch.rand(1000, generator=g))
    assert pos.numel() == 64 and neg.numel() == 192
    assert bool((lab[pos] == 1).all()) and bool((lab[neg] == 0).all())


def test_oracle_step_runs_and_decreases_nothing_weird():
    cfg = O.OracleCfg()
    s = O.OracleRCNN(cfg, seed=1)
    t = O.OracleRCNN(cfg, seed=1)
    opt = O.make_optimizer(s, cfg)
    H, W = 96, 128
    lq = O.synthetic_batch(1, H, W, 8, 1)
    uq = O.synthetic_batch(1, H, W, 8, 3, labelled=False)
    out = O.run_step(s, t, opt, (lq, [dict(d) for d in lq], uq, [dict(d) for d in uq]), cfg, [0.8], [0.9])
    assert len([k for k in out if k.startswith("loss")]) == 8
    assert all(math.isfinite(v) for v in out.values())


def test_pairwise_iou_matches_torchvision_and_zero_overlap_rule():
    """d2 v0.5 `pairwise_iou` = torchvision `box_iou` wherever boxes overlap and exactly 0 elsewhere (also for
    degenerate boxes, where inter / union would be 0 / 0)."""
    from torchvision.ops import box_iou
    g = torch.Generator().manual_seed(4)
    xy = torch.rand(40, 2, generator=g) * 200
    a = torch.cat([xy, xy + torch.rand(40, 2, generator=g) * 150 + 1], 1)
    xy = torch.rand(300, 2, generator=g) * 200
    b = torch.cat([xy, xy + torch.rand(300, 2, generator=g) * 150 + 1], 1)
    assert torch.allclose(O.pairwise_iou(a, b), box_iou(a, b), atol=1e-7)
    deg = torch.tensor([[5., 5., 5., 5.]])
    assert O.pairwise_iou(deg, deg).tolist() == [[0.0]]


def test_clip_and_nonempty():
    b = torch.tensor([[-3., -2., 50., 70.], [10., 10., 10., 30.], [20., 5., 25., 9.]])
    c = O.clip_boxes(b, (40, 30))  # (h, w): x in [0, 30], y in [0, 40]
    assert c.tolist() == [[0., 0., 30., 40.], [10., 10., 10., 30.], [20., 5., 25., 9.]]
    assert O.nonempty(c).tolist() == [True, False, True]
    assert O.nonempty(c, threshold=4.5).tolist() == [True, False, False]


def test_nms_ties_duplicates_and_exact_threshold():
    """Integer-grid boxes (IoUs exactly equal to the threshold), exact duplicates, zero-area boxes, five distinct
    scores: the oracle's NMS keeps exactly what torchvision keeps, in the same order; batched NMS keeps the same set."""
    from torchvision.ops import nms as tv_nms
    from torchvision.ops import boxes as tvb
    g = torch.Generator().manual_seed(0)
    exact = 0
    for trial in range(60):
        n = int(torch.randint(1, 400, (1,), generator=g))
        xy = torch.randint(0, 40, (n, 2), generator=g).float() * 4
        wh = torch.randint(0, 12, (n, 2), generator=g).float() * 8
        b = torch.cat([xy, xy + wh], 1)
        if trial % 3 == 0:
            b[::7] = b[0].clone()
        s = torch.randint(0, 5, (n,), generator=g).float() / 4
        for thr in (0.5, 0.7):
            assert torch.equal(O.nms(b, s, thr), tv_nms(b, s, thr))
            exact += int((O.pairwise_iou(b, b) == thr).sum())
        idx = torch.randint(0, 8, (n,), generator=g)
        # batched NMS: same kept SET and same score sequence; the order among EQUAL scores is implementation-defined
        # (torchvision's final sort is not stable, the oracle's and the CUDA path's is: ascending index)
        mine, tv = O.batched_nms(b, s, idx, 0.5), tvb._batched_nms_vanilla(b, s, idx, 0.5)
        assert torch.equal(mine.sort().values, tv.sort().values) and torch.equal(s[mine], s[tv])
    assert exact > 100  # the family really contains IoU == threshold pairs
