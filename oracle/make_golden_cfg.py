"""Generates tests/golden/pt_reference_cfg_golden.json: the reference's OWN configuration files, parsed with
`yaml.safe_load` and merged along their `_BASE_` chain by the few lines below (independent of the package's config
loader): configs/Guassian-RCNN-VGG.yaml <- configs/pt/final_{c2f,k2c,s2c,c2b,c2f_0.02}.yaml, plus the command-line
overrides of train.sh. tests/test_host_cpu.py checks `config.c2f_config()` / `k2c_config()` (the values the bench and
the tests run with) against it key by key.

    python oracle/make_golden_cfg.py

Test infrastructure: runs only here (the reference tree does not exist on the GPU box); the fixture is committed."""
import json
import os
import re
import sys

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PT_REFERENCE", "/root/reference")


def load(path):
    d = yaml.safe_load(open(path))
    base = d.pop("_BASE_", None)
    if base:
        merged = load(os.path.join(os.path.dirname(path), base))
        merge(merged, d)
        return merged
    return d


def merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            merge(dst[k], v)
        else:
            dst[k] = v


def train_sh_overrides():
    """KEY VALUE pairs after `--config ...` in train.sh."""
    text = open(os.path.join(REF, "train.sh")).read().replace("\\\n", " ")
    tail = text.split("--config", 1)[1].split()[1:]
    return {k: v for k, v in zip(tail[0::2], tail[1::2]) if re.match(r"^[A-Z_.]+$", k)}


def main():
    out = {"train_sh": train_sh_overrides()}
    for name in ("final_c2f", "final_k2c", "final_s2c", "final_c2b", "final_c2f_0.02"):
        out[name] = load(os.path.join(REF, "configs", "pt", name + ".yaml"))
    dst = os.path.join(os.environ.get("PT_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden")), "pt_reference_cfg_golden.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("train.sh:", out["train_sh"])
    print("wrote", dst)


if __name__ == "__main__":
    sys.exit(main())
