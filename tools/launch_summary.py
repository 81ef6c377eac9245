"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`):
per-kernel count, total time, share and - when the byte counters were collected - DRAM bytes per launch.

   python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches.md
   python tools/launch_summary.py gpurun_out/launches.csv --steps 3 --json profiles/r02_ncu_step_bytes.json --precision bf16

With --json the DRAM bytes of the conv-graph kernels (conv_* and fuse_sum*) are summed, divided by --steps (the number of
whole-path steps inside the captured window) and merged into the JSON file that bench.py reads for `roofline.traffic`."""
import argparse
import csv
import json
import os
import re
from collections import OrderedDict


def to_number(v):
    return float(v.replace(",", ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--steps", type=int, default=0)
    ap.add_argument("--json", type=str, default=None)
    ap.add_argument("--precision", type=str, default="bf16")
    ap.add_argument("--command", type=str, default="")
    args = ap.parse_args()
    with open(args.path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ii, ki, mi, vi, ui = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = OrderedDict()          # id -> [name, us, read bytes, write bytes]
    for r in rd:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("b200romp::", "").replace("void ", "")
        L = launches.setdefault(r[ii], [name, 0.0, 0.0, 0.0])
        v, u = to_number(r[vi]), r[ui].lower()
        if r[mi] == "gpu__time_duration.sum":
            L[1] = v / 1000.0 if u.startswith("ns") else (v if u.startswith("us") else v * 1000.0 if u.startswith("ms") else v * 1e6)
        elif r[mi] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
            L[2 if r[mi].endswith("read.sum") else 3] = v * scale
    agg = OrderedDict()
    for name, us, rb, wb in launches.values():
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += us; a[2] += rb; a[3] += wb
    total = sum(a[1] for a in agg.values())
    have_bytes = any(a[2] + a[3] > 0 for a in agg.values())
    if have_bytes:
        print("| kernel | launches | total us | share | avg us | DRAM read MB / launch | DRAM write MB / launch | GB/s |\n|---|---:|---:|---:|---:|---:|---:|---:|")
    else:
        print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for n, (c, t, rb, wb) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        row = f"| `{n}` | {c} | {t:.1f} | {100 * t / total:.1f}% | {t / c:.1f} |"
        if have_bytes:
            row += f" {rb / c / 1e6:.1f} | {wb / c / 1e6:.1f} | {(rb + wb) / t / 1e3:.0f} |"
        print(row)
    print(f"| **total** | {len(launches)} | {total:.1f} | 100% | |" + (" | | |" if have_bytes else ""))
    if args.json and have_bytes and args.steps > 0:
        conv = [(c, t, rb, wb) for n, (c, t, rb, wb) in agg.items() if n.startswith(("conv_", "fuse_sum"))]
        d = json.load(open(args.json)) if os.path.exists(args.json) else {}
        d[args.precision] = {
            "conv_graph_dram_bytes_per_step": sum(x[2] + x[3] for x in conv) / args.steps,
            "conv_graph_dram_read_bytes_per_step": sum(x[2] for x in conv) / args.steps,
            "conv_graph_dram_write_bytes_per_step": sum(x[3] for x in conv) / args.steps,
            "conv_graph_kernel_us_per_step_under_ncu": sum(x[1] for x in conv) / args.steps,
            "conv_graph_launches_per_step": sum(x[0] for x in conv) / args.steps,
            "all_kernels_dram_bytes_per_step": sum(a[2] + a[3] for a in agg.values()) / args.steps,
            "steps_in_window": args.steps, "batch": 64,
            "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none (serialised, "
                      "cold-cache launches) of: " + args.command}
        json.dump(d, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
