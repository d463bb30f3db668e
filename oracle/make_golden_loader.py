"""Generates tests/golden/pt_reference_loader_golden.json: the batches the REFERENCE'S OWN
`AspectRatioGroupedSemiSupDatasetTwoCrop` (pt/data/common.py:106-180, imported unmodified; its detectron2 base class is
answered by the permissive stub oracle/d2shim_any.py -- the subclass overrides both `__init__` and `__iter__`) emits
for seeded streams of (strong, weak) pairs with mixed orientations and several batch-size pairs.

    python oracle/make_golden_loader.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle import d2shim, d2shim_any  # noqa: E402

d2shim.install()      # registers the detectron2.* stub packages d2shim_any extends
d2shim_any.install()
from pt.data.common import AspectRatioGroupedSemiSupDatasetTwoCrop  # noqa: E402


def stream(seed, n, tag, p_wide):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        wide = rng.random() < p_wide
        w, h = (rng.randint(600, 1333), rng.randint(300, 599)) if wide else (rng.randint(300, 599), rng.randint(600, 1333))
        if rng.random() < 0.05:
            w = h  # square: goes with the "tall" group (w > h is False)
        out.append([{"id": f"{tag}{i}s", "width": w, "height": h}, {"id": f"{tag}{i}w", "width": w, "height": h}])
    return out


def main():
    cases = []
    for seed, n, bl, bu, p in ((1, 60, 2, 2, 0.7), (2, 80, 4, 2, 0.5), (3, 50, 1, 3, 0.9), (4, 40, 2, 2, 1.0)):
        lab, unl = stream(seed, n, "L", p), stream(seed + 100, n, "U", 1.0 - p if p < 1.0 else 1.0)
        ds = AspectRatioGroupedSemiSupDatasetTwoCrop((lab, unl), (bl, bu))
        batches = [[[d["id"] for d in part] for part in b] for b in ds]
        print(f"seed {seed}: {n} pairs, batch ({bl}, {bu}) -> {len(batches)} batches")
        cases.append(dict(label=lab, unlabel=unl, batch=[bl, bu], batches=batches))
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_loader_golden.json")
    json.dump(cases, open(dst, "w"))
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
