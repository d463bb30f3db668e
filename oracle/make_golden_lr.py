"""Generates tests/golden/pt_reference_lr_golden.json: learning rates produced by the REFERENCE'S OWN
`WarmupTwoStageMultiStepLR` (pt/solver/lr_scheduler.py:21-66, imported unmodified) stepping a torch SGD optimizer
once per iteration, for a few configurations. Its one detectron2 import, `_get_warmup_factor_at_iter`, is restated
below from detectron2 v0.5 (solver/lr_scheduler.py).

    python oracle/make_golden_lr.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed."""
import json
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def _get_warmup_factor_at_iter(method, iter, warmup_iters, warmup_factor):  # noqa: A002  (d2 v0.5 signature)
    if iter >= warmup_iters:
        return 1.0
    if method == "constant":
        return warmup_factor
    elif method == "linear":
        alpha = iter / warmup_iters
        return warmup_factor * (1 - alpha) + alpha
    raise ValueError("Unknown warmup method: {}".format(method))


for name in ("detectron2", "detectron2.solver", "detectron2.solver.lr_scheduler"):
    m = types.ModuleType(name)
    m.__path__ = []
    sys.modules.setdefault(name, m)
sys.modules["detectron2.solver.lr_scheduler"]._get_warmup_factor_at_iter = _get_warmup_factor_at_iter

import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_lr_scheduler", os.path.join(REF, "pt", "solver", "lr_scheduler.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [dict(base_lr=0.016, steps=[30, 60], factor_list=[1.0, 0.5, 0.05], warmup_factor=0.001, warmup_iters=10,
              warmup_method="linear", n=80),
         dict(base_lr=0.04, steps=[5], factor_list=[1, 10], warmup_factor=0.25, warmup_iters=8,
              warmup_method="constant", n=20),
         dict(base_lr=0.01, steps=[], factor_list=[2.0], warmup_factor=0.001, warmup_iters=0,
              warmup_method="linear", n=5)]


def main():
    out = []
    for c in CASES:
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=c["base_lr"])
        sch = ref.WarmupTwoStageMultiStepLR(opt, c["steps"], c["factor_list"], warmup_factor=c["warmup_factor"],
                                            warmup_iters=c["warmup_iters"], warmup_method=c["warmup_method"])
        lrs = []
        for _ in range(c["n"]):
            lrs.append(opt.param_groups[0]["lr"])  # the rate iteration `it` trains with
            opt.step()
            sch.step()
        out.append(dict(c, lrs=lrs))
        print(c["steps"], c["factor_list"], [round(x, 6) for x in lrs[:3]], "...", round(lrs[-1], 6))
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_lr_golden.json")
    json.dump(out, open(dst, "w"))
    print("wrote", dst)


if __name__ == "__main__":
    main()
