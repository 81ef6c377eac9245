"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.
   python tools/launch_summary.py gpurun_out/launches.csv > profiles/r01_launches.md"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1000.0 if u.startswith("ns") else (v if u.startswith("us") else v * 1000.0)
        name = re.sub(r"\(.*", "", r[ki]).replace("b200romp::", "").replace("void ", "")
        rows.append((name, us))
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t:.1f} | {100 * t / total:.1f}% | {t / c:.1f} |")
    print(f"| **total** | {len(rows)} | {total:.1f} | 100% | |")


if __name__ == "__main__":
    main(sys.argv[1])
