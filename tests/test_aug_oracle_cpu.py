"""CPU: the strong-augmentation oracle (oracle/aug_oracle.py) against (a) the REFERENCE'S OWN pipeline
(tests/golden/pt_reference_aug_golden.pt, made by oracle/make_golden_aug.py from the unmodified
pt/data/detection_utils.py + pt/data/transforms/augmentation_impl.py on this image's torchvision / Pillow) and (b) the
installed Pillow directly: colour-space conversions exhaustively over all 2^24 inputs, Gaussian blur over a radius sweep."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import aug_oracle as A

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "pt_reference_aug_golden.pt"), weights_only=False)


def test_pipeline_reproduces_the_reference_augmentation_bit_for_bit():
    flags = {"jitter": 0, "gray": 0, "blur": 0, "solarize": 0}
    for c in G["cases"]:
        torch.manual_seed(G["torch_seed_base"] + c["seed"])
        random.seed(G["py_seed_base"] + c["seed"])
        p = A.sample_params()
        out = A.strong_augment(c["image"].numpy(), p)
        assert np.array_equal(out, c["output"].numpy()), (c["seed"], vars(p))
        for k in flags:
            flags[k] += int(getattr(p, k))
    assert all(v >= 3 for v in flags.values()), flags  # every branch of the Compose is exercised


def test_colour_space_conversions_exhaustively_vs_pillow():
    from PIL import Image
    r, g, b = np.meshgrid(*(np.arange(256, dtype=np.uint8),) * 3, indexing="ij")
    cube = np.stack([r, g, b], -1).reshape(4096, 4096, 3)
    assert np.array_equal(A.rgb_to_hsv(cube), np.array(Image.fromarray(cube, "RGB").convert("HSV")))
    assert np.array_equal(A.hsv_to_rgb(cube), np.array(Image.fromarray(cube, "HSV").convert("RGB")))
    assert np.array_equal(A.to_gray(cube), np.array(Image.fromarray(cube, "RGB").convert("L")))


@pytest.mark.parametrize("shape", [(97, 131), (5, 3), (1, 40)])
def test_ops_vs_pillow_and_torchvision(shape):
    from PIL import Image, ImageEnhance, ImageFilter, ImageOps
    import torchvision.transforms.functional as F
    rs = np.random.RandomState(shape[0])
    img = rs.randint(0, 256, shape + (3,)).astype(np.uint8)
    im = Image.fromarray(img, "RGB")
    for sig in list(np.linspace(0.1, 2.0, 20)) + [3.3]:
        assert np.array_equal(A.gaussian_blur(img, float(sig)), np.array(im.filter(ImageFilter.GaussianBlur(radius=float(sig))))), sig
    for f in (0.6, 0.77, 1.0, 1.23, 1.4):
        assert np.array_equal(A.adjust_brightness(img, f), np.array(ImageEnhance.Brightness(im).enhance(f)))
        assert np.array_equal(A.adjust_contrast(img, f), np.array(ImageEnhance.Contrast(im).enhance(f)))
        assert np.array_equal(A.adjust_saturation(img, f), np.array(ImageEnhance.Color(im).enhance(f)))
    for h in (-0.1, -0.033, 0.0, 0.05, 0.1):
        assert np.array_equal(A.adjust_hue(img, h), np.array(F.adjust_hue(im, h)))
    assert np.array_equal(A.rgb_to_grayscale3(img), np.array(F.rgb_to_grayscale(im, 3)))
    assert np.array_equal(A.solarize(img), np.array(ImageOps.solarize(im, 128)))


# ------------------------------------------------------------------------------------------- weak augmentation
@pytest.mark.parametrize("hw,new", [((90, 60), (45, 30)), ((53, 106), (106, 212)), ((64, 33), (33, 64)), ((47, 90), (47, 45)),
                                    ((120, 200), (75, 125)), ((375, 500), (600, 800)), ((1, 7), (3, 2))])
def test_bilinear_resize_vs_pillow(hw, new):
    """d2's ResizeTransform.apply_image on uint8 = PIL Image.resize((w, h), BILINEAR) (libImaging/Resample.c)."""
    from PIL import Image
    rs = np.random.RandomState(hw[0] * 1000 + hw[1])
    img = rs.randint(0, 256, hw + (3,)).astype(np.uint8)
    want = np.asarray(Image.fromarray(img).resize((new[1], new[0]), Image.BILINEAR))
    assert np.array_equal(A.resize_bilinear(img, new[0], new[1]), want)


def test_shortest_edge_rule_and_box_transform():
    """ResizeShortestEdge.get_transform of d2 v0.5 (the reference pins detectron2 v0.5: README 'Installation') and the
    box side of ResizeTransform / HFlipTransform + the clip of transform_instance_annotations."""
    assert A.shortest_edge_size(1024, 2048, 600, 1200) == (600, 1200)
    assert A.shortest_edge_size(375, 500, 600, 1200) == (600, 800)
    assert A.shortest_edge_size(500, 375, 600, 1200) == (800, 600)
    assert A.shortest_edge_size(300, 1000, 600, 1200) == (360, 1200)
    assert A.shortest_edge_size(333, 500, 800, 1333) == (800, 1201)
    from PIL import Image
    rs = np.random.RandomState(5)
    img = rs.randint(0, 256, (40, 60, 3)).astype(np.uint8)
    boxes = np.array([[5.0, 3.0, 25.0, 21.0], [0.0, 0.0, 60.0, 40.0]])
    out, b = A.weak_augment(img, boxes, 50, 70, True)
    want = np.asarray(Image.fromarray(img).resize((70, 47), Image.BILINEAR))[:, ::-1]
    assert np.array_equal(out, want)
    s = 70 / 60
    assert np.allclose(b[0], [70 - 25 * s, 3 * 47 / 40, 70 - 5 * s, 21 * 47 / 40])
    assert np.allclose(b[1], [0, 0, 70, 47])
    out2, b2 = A.weak_augment(img, boxes, 50, 70, False)
    assert np.array_equal(out2, want[:, ::-1]) and np.allclose(b2[0], [5 * s, 3 * 47 / 40, 25 * s, 21 * 47 / 40])


def test_product_resample_tables_equal_the_oracle():
    from probabilisticteacher_b200 import data_aug as D
    for a, b in [(90, 60), (53, 106), (1333, 800), (64, 33), (47, 90), (80, 80), (2048, 1200), (3, 1)]:
        bo, ko = A.resample_coeffs(a, b)
        bn, kn, ks = D.resample_coeffs(a, b)
        assert np.array_equal(bo, bn) and np.array_equal(ko, kn) and ks == ko.shape[1]
        assert D.shortest_edge_size(a, b, 600, 1200) == A.shortest_edge_size(a, b, 600, 1200)
