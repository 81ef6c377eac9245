"""Per-op device-time table of the ROMP conv net (b200romp_net_profile): which layers the step is spent in.

    python tools/op_profile.py [--batch 64] [--iters 5] > profiles/rNN_op_profile.md

Ops are launched one by one (no CUDA graph) with a CUDA event between them, so the numbers are warm-cache device
times in network order; FLOP counts are algorithmic (2*MAC of the conv at its own resolution)."""
import argparse
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--top", type=int, default=400)
    ap.add_argument("--precision", type=str, default="bf16", choices=["bf16", "tf32", "fp32"])
    args = ap.parse_args()
    from romp_b200 import graph, synth, _lib

    torch.cuda.set_device(0)
    B = args.batch
    nb, io = graph.build_romp(synth.romp_state_dict(0), 0, args.precision, _lib.U8, max_batch=B)
    lib = nb.lib
    frames = torch.randint(0, 256, (B, 512, 512, 3), dtype=torch.uint8, device="cuda")
    ext = {}
    if io is not None:
        ext["frames"] = frames
        ext["center_maps"] = torch.empty(B, 1, 64, 64, device="cuda")
        ext["params_maps"] = torch.empty(B, 145, 64, 64, device="cuda")
        for k, t in ext.items():
            _lib.check(lib.b200romp_net_bind(nb.net, io[k], t.data_ptr()))
    us = nb.profile(B, args.iters)
    lines = [l for l in nb.describe().splitlines() if l.startswith("op")]
    rows = []
    for l, t in zip(lines, us):
        if " sum " in l:
            rows.append((t, 0.0, l))
            continue
        m = re.search(r"k(\d+) s(\d) +(\d+)->(\d+) +in t\d+\[(\d+)x(\d+)x\d+\]", l)
        k, s, cin, cout, H, W = (int(x) for x in m.groups())
        k2 = 3 if k == 13 else k * k
        ho, wo = H // s, W // s
        flop = 2.0 * B * ho * wo * cin * cout * k2
        rows.append((t, flop, l))
    total = sum(r[0] for r in rows)
    tf = sum(r[1] for r in rows)
    print(f"# per-op profile, batch {B}: {len(rows)} ops, {total / 1000:.3f} ms, {tf / total / 1e6:.1f} TFLOP/s overall\n")
    # by class
    cls = {}
    for t, f, l in rows:
        if " sum " in l:
            ms = re.search(r"out t\d+\[(\d+)x\d+x(\d+)\]", l)
            a = cls.setdefault(f"sum @{ms.group(1)} c{ms.group(2)} terms{l.count('up')}", [0, 0.0, 0.0])
            a[0] += 1; a[1] += t
            continue
        m = re.search(r"(tcgen05|simt) +k(\d+) s(\d) +(\d+)->(\d+) +in t\d+\[(\d+)x", l)
        key = f"{m.group(1)} k{m.group(2)} s{m.group(3)} {m.group(4)}->{m.group(5)} @{m.group(6)}" + (" epi" + l.split("epi")[1][0] if "epi" in l else "") + \
              (" up" + re.search(r"up(\d)", l).group(1))
        a = cls.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += f
    print("| class | ops | total us | share | avg us | TFLOP/s |\n|---|---:|---:|---:|---:|---:|")
    for k, (n, t, f) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {t:.1f} | {100 * t / total:.1f}% | {t / n:.1f} | {f / t / 1e6:.0f} |")
    print("\n| us | TFLOP/s | op |\n|---:|---:|---|")
    for t, f, l in sorted(rows, key=lambda r: -r[0])[:args.top]:
        print(f"| {t:.1f} | {f / t / 1e6:.0f} | `{l.strip()}` |")


if __name__ == "__main__":
    main()
