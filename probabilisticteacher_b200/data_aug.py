"""Strong augmentation of the input pipeline ON THE DEVICE: the counterpart of `pt/data/detection_utils.py:38-60`
(`build_strong_augmentation`: RandomApply(ColorJitter(0.4, 0.4, 0.4, 0.1), p=0.8), RandomGrayscale(p=0.2),
RandomApply(GaussianBlur([0.1, 2.0]), p=0.5), RandomApply(Solarize(0.5), p=0.2)) as `DatasetMapperTwoCropSeparate`
applies it to the weakly augmented image (`pt/data/dataset_mapper.py:159-172`). The reference runs torchvision /
Pillow on the CPU per image; here the uint8 CHW image stays in HBM and the chain runs as a handful of byte kernels
(csrc/augment.cu), bit-exact with Pillow 12.2 / torchvision 0.26.

The random decisions are drawn on the host IN TORCHVISION'S ORDER from the same generators (torch's global generator
or an explicit one; python's `random` for the blur sigma), so that under the same seeds this pipeline and the
reference's Compose produce the same image (tests/test_aug_gpu.py)."""
import random

import numpy as np
import torch

from ._lib import call

OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_HUE, OP_GRAY, OP_SOLARIZE = range(6)


class StrongAugParams:
    __slots__ = ("jitter", "order", "brightness", "contrast", "saturation", "hue", "gray", "blur", "sigma", "solarize")

    def __init__(self):
        self.jitter = False
        self.order = (0, 1, 2, 3)
        self.brightness = self.contrast = self.saturation = 1.0
        self.hue = 0.0
        self.gray = self.blur = self.solarize = False
        self.sigma = 0.0


def gaussian_box(radius, passes=3):
    """Pillow's box-blur approximation of a Gaussian of sigma = `radius` (libImaging/BoxBlur.c
    `_gaussian_blur_radius` + the 8.24 fixed-point tap weights of `ImagingHorizontalBoxBlur`), evaluated with the
    single-precision variables of the C code. Returns (integer radius, ww, fw)."""
    f32 = np.float32
    r = f32(radius)
    sigma2 = f32(f32(r * r) / f32(passes))
    L = f32(np.sqrt(12.0 * float(sigma2) + 1.0))
    l = f32(np.floor((float(L) - 1.0) / 2.0))
    a = f32(f32(f32(2) * l + f32(1)) * f32(f32(l * f32(l + f32(1))) - f32(f32(3) * sigma2)))
    a = f32(a / f32(f32(6) * f32(sigma2 - f32(f32(l + f32(1)) * f32(l + f32(1))))))
    fr = f32(l + a)
    rad = int(fr)
    ww = int(f32(1 << 24) / f32(fr * f32(2) + f32(1)))
    fw = ((1 << 24) - (rad * 2 + 1) * ww) // 2
    return float(fr), rad, ww, fw


class StrongAugmentation:
    """`build_strong_augmentation(cfg, is_train)` for device images: callable on a uint8 CUDA tensor [3, H, W]
    (the layout `dataset_dict["image"]` has, dataset_mapper.py:165-167), returns a new tensor."""

    def __init__(self, is_train=True, brightness=0.4, contrast=0.4, saturation=0.4, hue=0.1, p_jitter=0.8, p_gray=0.2,
                 p_blur=0.5, sigma=(0.1, 2.0), p_solarize=0.2, generator=None, py_random=random):
        self.is_train = is_train
        self.b, self.c, self.s, self.h = brightness, contrast, saturation, hue
        self.p_jitter, self.p_gray, self.p_blur, self.p_solarize = p_jitter, p_gray, p_blur, p_solarize
        self.sigma = sigma
        self.generator = generator
        self.py_random = py_random

    def sample(self):
        """One draw of the Compose's decisions in torchvision's order: RandomApply -> torch.rand(1);
        ColorJitter.get_params -> torch.randperm(4), then one uniform_ per factor; RandomGrayscale -> torch.rand(1);
        RandomApply(GaussianBlur) -> torch.rand(1), sigma from python's random.uniform; RandomApply(Solarize)."""
        g = self.generator
        p = StrongAugParams()
        if not self.is_train:
            return p
        p.jitter = not (self.p_jitter < float(torch.rand(1, generator=g)))
        if p.jitter:
            p.order = tuple(int(x) for x in torch.randperm(4, generator=g))
            p.brightness = float(torch.empty(1).uniform_(max(0, 1 - self.b), 1 + self.b, generator=g))
            p.contrast = float(torch.empty(1).uniform_(max(0, 1 - self.c), 1 + self.c, generator=g))
            p.saturation = float(torch.empty(1).uniform_(max(0, 1 - self.s), 1 + self.s, generator=g))
            p.hue = float(torch.empty(1).uniform_(-self.h, self.h, generator=g))
        p.gray = float(torch.rand(1, generator=g)) < self.p_gray
        p.blur = not (self.p_blur < float(torch.rand(1, generator=g)))
        if p.blur:
            p.sigma = self.py_random.uniform(self.sigma[0], self.sigma[1])
        p.solarize = not (self.p_solarize < float(torch.rand(1, generator=g)))
        return p

    @staticmethod
    def apply(image, p):
        """Applies the decisions `p` to a uint8 CUDA image [3, H, W]."""
        if not image.is_cuda or image.dtype != torch.uint8 or image.dim() != 3 or image.shape[0] != 3:
            raise ValueError("StrongAugmentation expects a uint8 CUDA tensor [3, H, W] (no CPU path)")
        _, H, W = image.shape
        cur = image.contiguous()
        runs = [[]]  # runs of per-pixel ops; a contrast op opens a new run (it needs the mean luma at that point)
        if p.jitter:
            for fn in p.order:
                if fn == 0:
                    runs[-1].append((OP_BRIGHTNESS, p.brightness))
                elif fn == 1:
                    runs.append([(OP_CONTRAST, p.contrast)])
                elif fn == 2:
                    runs[-1].append((OP_SATURATION, p.saturation))
                else:
                    runs[-1].append((OP_HUE, float(int(np.array(p.hue * 255).astype(np.uint8)))))
        if p.gray:
            runs[-1].append((OP_GRAY, 0.0))
        tail = [(OP_SOLARIZE, 0.0)] if p.solarize else []
        if not p.blur:
            runs[-1] += tail
            tail = []
        gsum = torch.zeros(2, dtype=torch.int64, device=image.device)
        slot = 0
        for k, run in enumerate(runs):
            needs_out = k + 1 < len(runs)
            if not run and not needs_out:
                continue
            out = torch.empty_like(cur)
            call("ptb200_aug_pointwise_u8", cur, out, H, W, len(run), [o for o, _ in run] or [0],
                 [float(f) for _, f in run] or [0.0], gsum[slot:] if run and run[0][0] == OP_CONTRAST else None,
                 gsum[1 - slot:] if needs_out else None)
            slot = 1 - slot
            cur = out
        if p.blur:
            fr, rad, ww, fw = gaussian_box(p.sigma)
            if fr != 0.0:
                tmp, out = torch.empty_like(cur), torch.empty_like(cur)
                call("ptb200_aug_boxblur_u8", cur, tmp, out, 3, H, W, rad, ww, fw, 3)
                cur = out
            if tail:
                out = torch.empty_like(cur)
                call("ptb200_aug_pointwise_u8", cur, out, H, W, 1, [OP_SOLARIZE], [0.0], None, None)
                cur = out
        return cur.clone() if cur is image else cur

    def __call__(self, image):
        return self.apply(image, self.sample())


def build_strong_augmentation(cfg, is_train, **kw):
    """Same name / arguments as `pt/data/detection_utils.py:38`."""
    return StrongAugmentation(is_train=is_train, **kw)


def two_crop(image_weak, strong_augmentation):
    """`DatasetMapperTwoCropSeparate.__call__`'s last step (dataset_mapper.py:159-172) for a device image: returns
    (strongly augmented image, weakly augmented image), both uint8 [3, H, W]."""
    return strong_augmentation(image_weak), image_weak


# ------------------------------------------------------------------------------------------ weak augmentation
# d2 v0.5 `utils.build_augmentation(cfg, is_train)` = [ResizeShortestEdge(MIN_SIZE_TRAIN, MAX_SIZE_TRAIN, sampling),
# RandomFlip(horizontal)] as the reference's mapper uses it (pt/data/dataset_mapper.py:67,104-106). The image ops run
# on the device, bit-exact with Pillow's bilinear `Image.resize`; sizes, decisions and the box transform stay on the host.
_PRECISION_BITS = 32 - 8 - 2
_coeff_cache = {}


def resample_coeffs(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (libImaging/Resample.c): int32
    bounds [out, 2] and 22-bit fixed-point coefficients [out, ksize]."""
    key = (in_size, out_size)
    hit = _coeff_cache.get(key)
    if hit is not None:
        return hit
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    ksize = int(np.ceil(support)) * 2 + 1
    xx = np.arange(out_size, dtype=np.float64)
    center = (xx + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    j = np.arange(ksize, dtype=np.float64)[None, :]
    a = np.abs((j + xmin[:, None] - center[:, None] + 0.5) * (1.0 / fscale))
    w = np.where((a < 1.0) & (j < xmax[:, None]), 1.0 - a, 0.0)
    ww = w.sum(1, keepdims=True)
    w = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    kk = np.where(w < 0, (-0.5 + w * (1 << _PRECISION_BITS)).astype(np.int64), (0.5 + w * (1 << _PRECISION_BITS)).astype(np.int64))
    kk = np.where(j < xmax[:, None], kk, 0)
    out = (np.stack([xmin, xmax], 1).astype(np.int32), kk.astype(np.int32), ksize)
    if len(_coeff_cache) < 64:
        _coeff_cache[key] = out
    return out


def resize_bilinear(image, new_h, new_w):
    """PIL `Image.resize((new_w, new_h), BILINEAR)` of a uint8 CUDA image [3, H, W]."""
    _, H, W = image.shape
    cur = image.contiguous()
    dev = image.device
    if new_w != W:
        b, k, ks = resample_coeffs(W, new_w)
        out = torch.empty(3, H, new_w, dtype=torch.uint8, device=dev)
        call("ptb200_aug_resample_u8", cur, out, 3, H, W, H, new_w, 1, torch.from_numpy(b).to(dev), torch.from_numpy(k).to(dev), ks)
        cur = out
    if new_h != H:
        b, k, ks = resample_coeffs(H, new_h)
        out = torch.empty(3, new_h, cur.shape[2], dtype=torch.uint8, device=dev)
        call("ptb200_aug_resample_u8", cur, out, 3, H, cur.shape[2], new_h, cur.shape[2], 0, torch.from_numpy(b).to(dev),
             torch.from_numpy(k).to(dev), ks)
        cur = out
    return cur


def shortest_edge_size(h, w, size, max_size):
    """d2 v0.5 ResizeShortestEdge.get_transform."""
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


class WeakAugmentation:
    """ResizeShortestEdge(short_edge_length, max_size, sample_style) + RandomFlip(prob, horizontal) on a uint8 CUDA
    image [3, H, W] and its XYXY boxes. Decisions come from numpy's global RNG in d2's order (np.random.choice /
    randint for the size, then np.random.uniform for the flip)."""

    def __init__(self, short_edge_length=(600,), max_size=1200, sample_style="choice", flip_prob=0.5, is_train=True):
        self.short_edge_length = tuple(short_edge_length)
        self.max_size = max_size
        self.is_range = sample_style == "range"
        self.flip_prob = flip_prob if is_train else 0.0

    def sample(self):
        if self.is_range:
            size = int(np.random.randint(self.short_edge_length[0], self.short_edge_length[1] + 1))
        else:
            size = int(np.random.choice(self.short_edge_length))
        flip = bool(np.random.uniform() < self.flip_prob) if self.flip_prob > 0 else False
        return size, flip

    def apply(self, image, boxes, size, flip):
        _, H, W = image.shape
        nh, nw = shortest_edge_size(H, W, size, self.max_size) if size != 0 else (H, W)
        out = resize_bilinear(image, nh, nw)
        b = np.asarray(boxes.detach().cpu() if torch.is_tensor(boxes) else boxes, dtype=np.float64).reshape(-1, 4).copy()
        b[:, 0::2] *= nw * 1.0 / W
        b[:, 1::2] *= nh * 1.0 / H
        if flip:
            flipped = torch.empty_like(out)
            call("ptb200_aug_hflip_u8", out, flipped, 3, nh, nw)
            out = flipped
            x1, x2 = nw - b[:, 2], nw - b[:, 0]
            b[:, 0], b[:, 2] = x1, x2
        b = np.minimum(b.clip(min=0), [nw, nh, nw, nh])   # transform_instance_annotations clips to the image
        return out, torch.from_numpy(b).to(torch.float32)

    def __call__(self, image, boxes):
        size, flip = self.sample()
        return self.apply(image, boxes, size, flip)


def map_two_crop(image, boxes, classes, weak, strong):
    """`DatasetMapperTwoCropSeparate.__call__` (pt/data/dataset_mapper.py:88-172) for an image that is already on the
    device: weak augmentation of image + boxes, then the strong augmentation of the weak image; returns the
    (strong, weak) pair of dataset dicts the paired loader consumes (both carry the same instances)."""
    from .structures import Boxes, FreeInstances
    img_w, b = weak(image, boxes)
    h, w = img_w.shape[-2:]
    keep = (b[:, 2] > b[:, 0]) & (b[:, 3] > b[:, 1])    # utils.filter_empty_instances
    inst = FreeInstances((h, w), gt_boxes=Boxes(b[keep]), gt_classes=torch.as_tensor(classes)[keep])
    strong_img = strong(img_w)
    return ({"image": strong_img, "height": h, "width": w, "instances": inst},
            {"image": img_w, "height": h, "width": w, "instances": inst})
