"""TEST INFRASTRUCTURE: CPU (numpy) restatement of the STRONG AUGMENTATION of the reference's input pipeline.

Reference call sites: `pt/data/detection_utils.py:38-60` (`build_strong_augmentation`: RandomApply(ColorJitter(0.4, 0.4,
0.4, 0.1), p=0.8), RandomGrayscale(p=0.2), RandomApply(GaussianBlur([0.1, 2.0]), p=0.5), RandomApply(Solarize(0.5),
p=0.2)), `pt/data/transforms/augmentation_impl.py:22-53` (GaussianBlur = PIL `ImageFilter.GaussianBlur(radius=sigma)`,
Solarize = PIL `ImageOps.solarize(img, round(0.5 * 256))`), applied to a PIL image at `pt/data/dataset_mapper.py:159-164`.

The arithmetic lives in third-party dependencies that are NOT vendored in the reference: torchvision's PIL back end
(`torchvision.transforms`, v0.26 here) and Pillow 12.2.0 (`ImageEnhance` / `Image.blend`, `Image.convert` L / HSV,
`ImageFilter.GaussianBlur` = three box-blur passes per axis, libImaging/BoxBlur.c, Convert.c, Blend.c). Their published
algorithms are restated here and PINNED against the installed Pillow / torchvision: the colour-space conversions
EXHAUSTIVELY (all 2^24 inputs), everything else on random images and parameter sweeps, the whole pipeline against
`build_strong_augmentation`'s torchvision Compose under identical seeds (tests/test_aug_oracle_cpu.py).

All functions take / return uint8 arrays [H, W, 3] (the channel order PIL is told is "RGB")."""
import math
import random

import numpy as np

f32, f64 = np.float32, np.float64


# ------------------------------------------------------------------------------------------ Pillow primitives
def to_gray(img):
    """Image.convert("L"): ITU-R 601-2 luma in 16.16 fixed point (libImaging/Convert.c, L24 macro)."""
    r, g, b = (img[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def blend(degenerate, img, factor):
    """Image.blend(degenerate, img, factor) as ImageEnhance uses it (libImaging/Blend.c): single-precision
    `d + alpha * (i - d)`, truncated to uint8 for alpha in [0, 1], clipped to [0, 255] then truncated otherwise."""
    a = f32(factor)
    d = degenerate.astype(f32)
    t = d + a * (img.astype(f32) - d)
    if 0.0 <= a <= 1.0:
        return t.astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t)).astype(np.uint8)


def adjust_brightness(img, factor):
    return blend(np.zeros_like(img), img, factor)


def gray_mean(img):
    """int(ImageStat.Stat(img.convert("L")).mean[0] + 0.5): the degenerate image of ImageEnhance.Contrast."""
    g = to_gray(img)
    return int(float(g.astype(np.int64).sum()) / g.size + 0.5)


def adjust_contrast(img, factor):
    return blend(np.full_like(img, gray_mean(img)), img, factor)


def adjust_saturation(img, factor):
    return blend(np.repeat(to_gray(img)[..., None], 3, 2), img, factor)


def rgb_to_hsv(img):
    """Image.convert("HSV") (Convert.c rgb2hsv_row, "following colorsys.py"): channel ratios in float, the branch value
    with double intermediates rounded to float, fmod(h / 6 + 1, 1) in double rounded to float, then (int)(h * 255.0)."""
    r, g, b = (img[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(np.maximum(r, g), b)
    minc = np.minimum(np.minimum(r, g), b)
    cr = (maxc - minc).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = cr / maxc.astype(f32)
        rc = ((maxc - r).astype(f32) / cr).astype(f64)
        gc = ((maxc - g).astype(f32) / cr).astype(f64)
        bc = ((maxc - b).astype(f32) / cr).astype(f64)
        h = np.where(r == maxc, (bc - gc).astype(f32),
                     np.where(g == maxc, (2.0 + rc - bc).astype(f32), (4.0 + gc - rc).astype(f32))).astype(f32)
        hh = np.fmod(h.astype(f64) / 6.0 + 1.0, 1.0).astype(f32).astype(f64)
        uh = (hh * 255.0).astype(np.int64)
        us = (s.astype(f64) * 255.0).astype(np.int64)
    same = minc == maxc
    uh = np.where(same, 0, np.clip(uh, 0, 255))
    us = np.where(same, 0, np.clip(us, 0, 255))
    return np.stack([uh, us, maxc], -1).astype(np.uint8)


def _cround(x):
    """C round(): halves away from zero."""
    return np.where(x >= 0, np.floor(x + 0.5), np.ceil(x - 0.5))


def hsv_to_rgb(hsv):
    """Image.convert("RGB") from "HSV" (Convert.c hsv2rgb): i = floor(h * 6 / 255), f and s / 255 stored as floats,
    p / q / t = round(v * (1 - ...)) in double."""
    h = hsv[..., 0].astype(f32).astype(f64)
    s = hsv[..., 1]
    vv = hsv[..., 2]
    v = vv.astype(f32).astype(f64)
    i = np.floor(h * 6.0 / 255.0)
    f = (h * 6.0 / 255.0 - i.astype(f32).astype(f64)).astype(f32).astype(f64)
    fs = (s.astype(f32).astype(f64) / 255.0).astype(f32).astype(f64)
    p = np.clip(_cround(v * (1.0 - fs)), 0, 255).astype(np.uint8)
    q = np.clip(_cround(v * (1.0 - fs * f)), 0, 255).astype(np.uint8)
    t = np.clip(_cround(v * (1.0 - fs * (1.0 - f))), 0, 255).astype(np.uint8)
    k = i.astype(np.int64) % 6
    out = np.stack([np.choose(k, [vv, q, p, p, t, vv]), np.choose(k, [t, vv, vv, q, p, p]),
                    np.choose(k, [p, p, t, vv, vv, q])], -1)
    gray = s == 0
    out[gray] = np.repeat(vv[gray][:, None], 3, 1)
    return out


def adjust_hue(img, hue_factor):
    """torchvision F.adjust_hue on a PIL image (transforms/_functional_pil.py): H channel += uint8(hue_factor * 255)
    with uint8 wrap-around."""
    hsv = rgb_to_hsv(img)
    shift = int(np.array(hue_factor * 255).astype(np.uint8))
    hsv[..., 0] = (hsv[..., 0].astype(np.int64) + shift).astype(np.uint8)
    return hsv_to_rgb(hsv)


def rgb_to_grayscale3(img):
    """RandomGrayscale -> F.rgb_to_grayscale(img, 3): the L image replicated into three channels."""
    return np.repeat(to_gray(img)[..., None], 3, 2)


def solarize(img, threshold=128):
    """ImageOps.solarize: values >= threshold are inverted."""
    return np.where(img < threshold, img, 255 - img).astype(np.uint8)


def gaussian_box_radius(radius, passes=3):
    """BoxBlur.c _gaussian_blur_radius: the (fractional) box radius whose `passes`-fold box blur approximates a
    Gaussian of sigma = radius (Gwosdek et al.); single-precision variables, double-precision literals."""
    r = f32(radius)
    sigma2 = f32(f32(r * r) / f32(passes))
    L = f32(math.sqrt(12.0 * float(sigma2) + 1.0))
    l = f32(math.floor((float(L) - 1.0) / 2.0))
    a = f32(f32(f32(2) * l + f32(1)) * f32(f32(l * f32(l + f32(1))) - f32(f32(3) * sigma2)))
    a = f32(a / f32(f32(6) * f32(sigma2 - f32(f32(l + f32(1)) * f32(l + f32(1))))))
    return f32(l + a)


def box_weights(float_radius):
    """(radius, ww, fw) of BoxBlur.c ImagingHorizontalBoxBlur: 8.24 fixed-point weights of the full box taps (ww) and
    of the two fractional outer taps (fw)."""
    fr = f32(float_radius)
    radius = int(fr)
    ww = int(f32(1 << 24) / f32(fr * f32(2) + f32(1)))
    fw = ((1 << 24) - (radius * 2 + 1) * ww) // 2
    return radius, ww, fw


def box_blur_1d(img, float_radius, axis):
    """One ImagingHorizontalBoxBlur pass along `axis` with edge extension:
    out[x] = (ww * sum_{|d| <= radius} in[clamp(x + d)] + fw * (in[clamp(x - radius - 1)] + in[clamp(x + radius + 1)])
              + 2^23) >> 24."""
    radius, ww, fw = box_weights(float_radius)
    a = np.moveaxis(img, axis, 0).astype(np.int64)
    n = a.shape[0]
    idx = np.arange(n)
    acc = np.zeros_like(a)
    for d in range(-radius, radius + 1):
        acc += a[np.clip(idx + d, 0, n - 1)]
    far = a[np.clip(idx - radius - 1, 0, n - 1)] + a[np.clip(idx + radius + 1, 0, n - 1)]
    out = ((acc * ww + far * fw + (1 << 23)) >> 24).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def gaussian_blur(img, radius, passes=3):
    """ImageFilter.GaussianBlur(radius): `passes` box blurs along x, then `passes` along y (each rounded to uint8)."""
    fr = gaussian_box_radius(radius, passes)
    out = img
    if float(fr) != 0.0:
        for _ in range(passes):
            out = box_blur_1d(out, fr, 1)
        for _ in range(passes):
            out = box_blur_1d(out, fr, 0)
    return out


# ------------------------------------------------------------------------------------------ the pipeline
class StrongAugParams:
    """One draw of `build_strong_augmentation`'s random decisions, in torchvision's draw order
    (RandomApply.forward: torch.rand(1); ColorJitter.get_params: torch.randperm(4) then one uniform_ per factor;
    RandomGrayscale: torch.rand(1); GaussianBlur's sigma: python `random.uniform`)."""

    def __init__(self):
        self.jitter = False
        self.order = (0, 1, 2, 3)
        self.brightness = self.contrast = self.saturation = 1.0
        self.hue = 0.0
        self.gray = False
        self.blur = False
        self.sigma = 0.0
        self.solarize = False


def sample_params(torch_generator=None, py_random=random, brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1,
                  p_jitter=0.8, p_gray=0.2, p_blur=0.5, sigma=(0.1, 2.0), p_solarize=0.2):
    import torch
    g = torch_generator
    p = StrongAugParams()
    p.jitter = not (p_jitter < float(torch.rand(1, generator=g)))          # RandomApply: skip if p < rand
    if p.jitter:
        p.order = tuple(int(x) for x in torch.randperm(4, generator=g))
        p.brightness = float(torch.empty(1).uniform_(max(0, 1 - brightness), 1 + brightness, generator=g))
        p.contrast = float(torch.empty(1).uniform_(max(0, 1 - contrast), 1 + contrast, generator=g))
        p.saturation = float(torch.empty(1).uniform_(max(0, 1 - saturation), 1 + saturation, generator=g))
        p.hue = float(torch.empty(1).uniform_(-hue, hue, generator=g))
    p.gray = float(torch.rand(1, generator=g)) < p_gray
    p.blur = not (p_blur < float(torch.rand(1, generator=g)))
    if p.blur:
        p.sigma = py_random.uniform(sigma[0], sigma[1])
    p.solarize = not (p_solarize < float(torch.rand(1, generator=g)))
    return p


def strong_augment(img, p):
    """The Compose of detection_utils.py:50-57 applied with the decisions `p`."""
    out = img
    if p.jitter:
        for fn_id in p.order:
            if fn_id == 0:
                out = adjust_brightness(out, p.brightness)
            elif fn_id == 1:
                out = adjust_contrast(out, p.contrast)
            elif fn_id == 2:
                out = adjust_saturation(out, p.saturation)
            else:
                out = adjust_hue(out, p.hue)
    if p.gray:
        out = rgb_to_grayscale3(out)
    if p.blur:
        out = gaussian_blur(out, p.sigma)
    if p.solarize:
        out = solarize(out, 128)
    return out


# ------------------------------------------------------------------------------------------ weak augmentation
# detectron2 v0.5 (un-vendored): `utils.build_augmentation` = ResizeShortestEdge(MIN_SIZE_TRAIN, MAX_SIZE_TRAIN,
# MIN_SIZE_TRAIN_SAMPLING) + RandomFlip(horizontal), used by the reference's mapper at pt/data/dataset_mapper.py:67,104-106.
# Their image arithmetic is Pillow's `Image.resize(..., BILINEAR)` (libImaging/Resample.c) and a column reversal; the
# resize is restated here and pinned to the installed Pillow (tests/test_aug_oracle_cpu.py); the size rule and the box
# transforms restate d2's data/transforms/{augmentation_impl,transform}.py.
_PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter over the whole axis:
    (bounds [out, 2] = (first input index, tap count), coefficients [out, ksize] in 22-bit fixed point)."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int64)
    ss = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - a if a < 1.0 else 0.0)
        ww = sum(w)
        if ww != 0.0:
            w = [v / ww for v in w]
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def resample_1d(img, out_size, axis):
    a = np.moveaxis(img, axis, 0).astype(np.int64)
    bounds, kk = resample_coeffs(a.shape[0], out_size)
    out = np.zeros((out_size,) + a.shape[1:], np.int64)
    for xx in range(out_size):
        xmin, xmax = bounds[xx]
        ss = np.full(a.shape[1:], 1 << (_PRECISION_BITS - 1), np.int64)
        for x in range(xmax):
            ss += a[xmin + x] * kk[xx, x]
        out[xx] = np.clip(ss >> _PRECISION_BITS, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 0, axis)


def resize_bilinear(img, new_h, new_w):
    """PIL Image.resize((new_w, new_h), BILINEAR) of a uint8 image: horizontal pass, then vertical pass, each with
    uint8 output."""
    out = img
    if new_w != img.shape[1]:
        out = resample_1d(out, new_w, 1)
    if new_h != img.shape[0]:
        out = resample_1d(out, new_h, 0)
    return out


def shortest_edge_size(h, w, size, max_size):
    """d2 v0.5 ResizeShortestEdge.get_transform: output (h, w) for a sampled short-edge `size`."""
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)


def weak_augment(img, boxes, size, max_size, flip):
    """ResizeShortestEdge + RandomFlip(horizontal) applied to an image [H, W, 3] and XYXY boxes [n, 4] (float64 as d2's
    apply_box, clipped to the image as transform_instance_annotations does)."""
    h, w = img.shape[:2]
    nh, nw = shortest_edge_size(h, w, size, max_size)
    out = resize_bilinear(img, nh, nw)
    b = np.asarray(boxes, dtype=np.float64).copy().reshape(-1, 4)
    b[:, 0::2] *= nw * 1.0 / w
    b[:, 1::2] *= nh * 1.0 / h
    if flip:
        out = out[:, ::-1].copy()
        x1 = nw - b[:, 2]
        x2 = nw - b[:, 0]
        b[:, 0], b[:, 2] = x1, x2
    b = np.minimum(b.clip(min=0), [nw, nh, nw, nh])
    return out, b
