"""GPU: the device-side strong augmentation (probabilisticteacher_b200/data_aug.py, csrc/augment.cu) -- BIT-EXACT
against (a) the reference's own pipeline (tests/golden/pt_reference_aug_golden.pt: `build_strong_augmentation` of
pt/data/detection_utils.py:38-60 run on torchvision / Pillow under recorded seeds) and (b) the CPU oracle
(oracle/aug_oracle.py, pinned to Pillow exhaustively in tests/test_aug_oracle_cpu.py) at full image size."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_aug_golden.pt"), weights_only=False)


def _chw(img_hwc, dev):
    return torch.from_numpy(np.ascontiguousarray(img_hwc.transpose(2, 0, 1))).to(dev)


def _hwc(t):
    return t.cpu().numpy().transpose(1, 2, 0)


def test_reference_pipeline_golden_bit_exact(cuda):
    from probabilisticteacher_b200.data_aug import build_strong_augmentation
    aug = build_strong_augmentation(None, True)
    for c in G["cases"]:
        torch.manual_seed(G["torch_seed_base"] + c["seed"])
        random.seed(G["py_seed_base"] + c["seed"])
        out = aug(_chw(c["image"].numpy(), cuda))     # draws its decisions in torchvision's order
        torch.cuda.synchronize()
        assert np.array_equal(_hwc(out), c["output"].numpy()), c["seed"]


@pytest.mark.parametrize("H,W", [(800, 1333), (37, 53), (1, 40), (600, 2000)])
def test_every_op_vs_oracle_at_full_size(cuda, H, W):
    from oracle import aug_oracle as A
    from probabilisticteacher_b200.data_aug import StrongAugmentation, StrongAugParams
    rs = np.random.RandomState(H)
    img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    dev_img = _chw(img, cuda)
    cases = []
    for order in ((0, 1, 2, 3), (3, 2, 1, 0), (1, 0, 3, 2), (2, 3, 0, 1)):
        p = StrongAugParams()
        p.jitter, p.order = True, order
        p.brightness, p.contrast, p.saturation, p.hue = 0.6 + 0.2 * order[0], 1.4 - 0.2 * order[1], 0.61 + 0.25 * order[2], -0.1 + 0.06 * order[3]
        cases.append(p)
    for sig in (0.1, 0.77, 1.3, 2.0):
        p = StrongAugParams()
        p.blur, p.sigma = True, sig
        cases.append(p)
    p = StrongAugParams()
    p.gray = p.solarize = p.blur = True
    p.sigma = 1.9
    cases.append(p)
    p = StrongAugParams()
    p.jitter, p.order, p.contrast, p.hue, p.solarize = True, (1, 3, 0, 2), 0.73, 0.1, True
    cases.append(p)
    for p in cases:
        q = A.StrongAugParams()
        for k in ("jitter", "order", "brightness", "contrast", "saturation", "hue", "gray", "blur", "sigma", "solarize"):
            setattr(q, k, getattr(p, k))
        want = A.strong_augment(img, q)
        got = _hwc(StrongAugmentation.apply(dev_img, p))
        torch.cuda.synchronize()
        assert np.array_equal(got, want), {k: getattr(p, k) for k in p.__slots__}
    assert np.array_equal(_hwc(dev_img), img)  # the input is not modified


def test_decision_sampling_matches_the_oracle_sampler(cuda):
    """Same draws from the same generators -> same decisions (torchvision's order)."""
    from oracle import aug_oracle as A
    from probabilisticteacher_b200.data_aug import StrongAugmentation
    aug = StrongAugmentation()
    for seed in range(50):
        torch.manual_seed(seed)
        random.seed(seed)
        a = aug.sample()
        torch.manual_seed(seed)
        random.seed(seed)
        b = A.sample_params()
        for k in a.__slots__:
            assert getattr(a, k) == getattr(b, k), (seed, k)
    with pytest.raises(ValueError):
        StrongAugmentation.apply(torch.zeros(3, 4, 4, dtype=torch.uint8), aug.sample())   # no CPU path


@pytest.mark.parametrize("H,W,size,max_size", [(1024, 2048, 600, 1200), (375, 500, 600, 1200), (500, 375, 800, 1333),
                                               (37, 53, 64, 100), (600, 1200, 600, 1200), (90, 60, 45, 1000)])
def test_weak_augmentation_bit_exact_vs_oracle(cuda, H, W, size, max_size):
    """ResizeShortestEdge + RandomFlip on the device == the oracle (pinned to Pillow's bilinear resize on the CPU)."""
    from oracle import aug_oracle as A
    from probabilisticteacher_b200.data_aug import WeakAugmentation
    rs = np.random.RandomState(H + W)
    img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    boxes = np.stack([rs.uniform(0, W / 2, 5), rs.uniform(0, H / 2, 5), rs.uniform(W / 2, W, 5), rs.uniform(H / 2, H, 5)], 1)
    aug = WeakAugmentation((size,), max_size)
    for flip in (False, True):
        want_img, want_b = A.weak_augment(img, boxes, size, max_size, flip)
        got, b = aug.apply(_chw(img, cuda), torch.from_numpy(boxes), size, flip)
        torch.cuda.synchronize()
        assert np.array_equal(_hwc(got), want_img)
        assert np.allclose(b.numpy(), want_b.astype(np.float32), rtol=0, atol=1e-4)


def test_two_crop_mapper_on_device(cuda):
    """The (strong, weak) pair of DatasetMapperTwoCropSeparate: same instances on both, strong = aug(weak image)."""
    from oracle import aug_oracle as A
    from probabilisticteacher_b200.data_aug import StrongAugmentation, WeakAugmentation, map_two_crop
    rs = np.random.RandomState(0)
    img = rs.randint(0, 256, (120, 200, 3)).astype(np.uint8)
    boxes = torch.tensor([[10.0, 20.0, 100.0, 90.0], [50.0, 5.0, 50.0, 80.0]])   # the second is empty -> filtered
    np.random.seed(3); torch.manual_seed(3); random.seed(3)
    strong, weak = map_two_crop(_chw(img, cuda), boxes, [1, 2], WeakAugmentation((60, 75), 110), StrongAugmentation())
    np.random.seed(3)
    size = int(np.random.choice((60, 75)))
    flip = bool(np.random.uniform() < 0.5)
    want_img, want_b = A.weak_augment(img, boxes.numpy().astype(np.float64), size, 110, flip)
    assert np.array_equal(_hwc(weak["image"]), want_img)
    assert len(weak["instances"]) == 1 and weak["instances"] is strong["instances"]
    assert np.allclose(weak["instances"].gt_boxes.tensor.cpu().numpy(), want_b[:1].astype(np.float32), atol=1e-4)
    torch.manual_seed(3); random.seed(3)
    p = A.sample_params()
    assert np.array_equal(_hwc(strong["image"]), A.strong_augment(want_img, p))


def test_device_resize_vs_pillow_directly(cuda):
    """The device resize against Pillow itself (the library d2's ResizeTransform calls for uint8 images), not only
    against the oracle's restatement of it."""
    Image = pytest.importorskip("PIL.Image")
    from probabilisticteacher_b200.data_aug import resize_bilinear
    rs = np.random.RandomState(11)
    for (h, w), (nh, nw) in [((375, 500), (600, 800)), ((1024, 2048), (600, 1200)), ((61, 47), (23, 90))]:
        img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        want = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BILINEAR))
        got = _hwc(resize_bilinear(_chw(img, cuda), nh, nw))
        assert np.array_equal(got, want), ((h, w), (nh, nw))
