"""Summaries of ncu CSV logs for profiles/ (per-kernel launches, time, DRAM bytes, tensor-pipe activity).

  python tools/summarize_ncu.py traffic <csv> <steps> <out.json>   # --metrics gpu__time_duration.sum,dram__bytes_*,sm__pipe_tensor_cycles_active...
"""
import collections
import csv
import json
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    per = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = per.setdefault(r["ID"], {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("ptb::", "").replace("(bool)", "")})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)  # -> us
        if r["Metric Name"].startswith("dram__bytes"):
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[r["Metric Name"]] = v
    return per


def traffic(path, steps, out):
    per = load(path)
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "tw": 0.0})
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["us"] += t
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
        a["tw"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    res = collections.OrderedDict()
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        if a["us"] / steps < 20:
            continue
        res[k] = {"launches_per_step": a["n"] / steps, "ms_per_step": a["us"] / steps / 1e3,
                  "dram_read_MB_per_launch": a["rd"] / a["n"] / 1e6, "dram_write_MB_per_launch": a["wr"] / a["n"] / 1e6,
                  "dram_GBps": (a["rd"] + a["wr"]) / a["us"] / 1e3,
                  "tensor_pipe_active_pct_time_weighted": a["tw"] / a["us"] if a["us"] else 0.0}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], float(sys.argv[3]), sys.argv[4])
