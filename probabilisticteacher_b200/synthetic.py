"""Synthetic inputs of the benchmark: uint8 uniform BGR images and, for source images,
12 ground-truth boxes (w, h ~ U(32, 400), inside the image, uniform classes) in the reference's input format
(list of dicts with "image" and "instances", pt/data/dataset_mapper.py:162-169). Draw order per image is
image -> (w, h) -> centre -> classes from one seeded CPU generator, the same order as the oracle's
generator, so that both sides can be fed identical batches."""
import torch

from .structures import Boxes, FreeInstances


def synthetic_batch(n, H, W, num_classes, seed, boxes_per_image=12, labelled=True):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n):
        img = torch.randint(0, 256, (3, H, W), generator=g, dtype=torch.uint8)
        d = {"image": img, "height": H, "width": W}
        if labelled:
            wh = torch.rand(boxes_per_image, 2, generator=g) * (400 - 32) + 32
            wh[:, 0].clamp_(max=W - 2)
            wh[:, 1].clamp_(max=H - 2)
            cxy = torch.rand(boxes_per_image, 2, generator=g)
            x1 = cxy[:, 0] * (W - wh[:, 0])
            y1 = cxy[:, 1] * (H - wh[:, 1])
            boxes = torch.stack([x1, y1, x1 + wh[:, 0], y1 + wh[:, 1]], 1)
            cls = torch.randint(0, num_classes, (boxes_per_image,), generator=g)
            d["instances"] = FreeInstances((H, W), gt_boxes=Boxes(boxes), gt_classes=cls)
        out.append(d)
    return out
