#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file x.csv <cmd>`):
launches, summed duration and share of all kernel time.  Times under ncu are serialised and cold-cache: the SHARES are
what is comparable with bench.py's own CUDA-event numbers.  usage: tools/ncu_launch_summary.py x.csv > summary.txt"""
import collections
import csv
import re
import sys


def main(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
    d = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        except ValueError:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"^(cub::\w+)::.*?(Device\w+Kernel).*", r"\1 \2", name)
        name = name if name.startswith("k_ydrop") else re.sub(r"<.*", "", name)
        d[name][0] += 1
        d[name][1] += v
    tot = sum(v[1] for v in d.values())
    print(f"{path}: {sum(v[0] for v in d.values())} launches, {tot:.1f} ms of kernel time (serialised under ncu)")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k[:70]:70s} n={v[0]:6d} {v[1]:11.2f} ms {100 * v[1] / tot:6.2f} %")


if __name__ == "__main__":
    main(sys.argv[1])
