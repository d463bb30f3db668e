"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, steps):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in rows:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else (v * 1e6 if unit == "s" else v))
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / steps:.1f} us/step over {steps} steps ({len(rows)} launches)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print(f"{v[1] / steps:10.1f} us/step {v[0] / steps:7.1f} launches/step {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
