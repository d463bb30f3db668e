"""Generates tests/golden/pt_reference_aug_golden.pt: outputs of the REFERENCE'S OWN strong augmentation
(`pt/data/detection_utils.py:38-60` `build_strong_augmentation`, with `pt/data/transforms/augmentation_impl.py`'s
GaussianBlur / Solarize, imported unmodified from /root/reference; torchvision + Pillow as installed here) on seeded
uint8 images, exactly as `DatasetMapperTwoCropSeparate.__call__` applies it (`pt/data/dataset_mapper.py:159-164`:
numpy -> PIL "RGB" -> Compose -> numpy). For every case the torch / python RNG seeds, the input image and the output
image are stored; tests/test_aug_oracle_cpu.py replays the same draws with oracle/aug_oracle.py.

    python oracle/make_golden_aug.py"""
import os
import random
import sys

import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("PT_REFERENCE", "/root/reference")
sys.path.insert(0, REF)

from oracle import d2shim, d2shim_any  # noqa: E402

d2shim.install()
d2shim_any.install()
from pt.data.detection_utils import build_strong_augmentation  # noqa: E402


def main():
    aug = build_strong_augmentation(None, True)
    cases = []
    rs = np.random.RandomState(7)
    sizes = [(37, 53), (48, 64), (64, 80), (33, 120), (72, 45)]
    for seed in range(32):
        h, w = sizes[seed % len(sizes)]
        kind = seed % 4
        if kind == 0:
            img = rs.randint(0, 256, (h, w, 3)).astype(np.uint8)
        elif kind == 1:  # smooth gradients (blur / hue see structure)
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // (h + w - 2))], -1).astype(np.uint8)
        elif kind == 2:  # dark, low-saturation
            img = (rs.randint(0, 40, (h, w, 1)) + rs.randint(0, 8, (h, w, 3))).astype(np.uint8)
        else:            # blocks of saturated colours
            img = np.kron(rs.randint(0, 2, ((h + 7) // 8, (w + 7) // 8, 3)) * 255, np.ones((8, 8, 1)))[:h, :w].astype(np.uint8)
        torch.manual_seed(1000 + seed)
        random.seed(2000 + seed)
        out = np.array(aug(Image.fromarray(img, "RGB")))
        cases.append(dict(seed=seed, image=torch.from_numpy(img), output=torch.from_numpy(out)))
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_aug_golden.pt")
    torch.save(dict(cases=cases, torch_seed_base=1000, py_seed_base=2000), dst)
    changed = sum(int(not torch.equal(c["image"], c["output"])) for c in cases)
    print("wrote", dst, os.path.getsize(dst), "bytes;", changed, "of", len(cases), "images changed")


if __name__ == "__main__":
    main()
