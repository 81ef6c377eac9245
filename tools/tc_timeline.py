"""Phase timeline of the CTA-pair conv kernels inside a replayed CUDA graph (B200ROMP_TC_STAMPS=1) and the per-op time as a
function of the batch (fixed cost per launch vs cost per tile).

    python tools/tc_timeline.py --cin 128 --hw 32      # a chain of 6 identical 3x3 convs with residuals, like an HRNet branch
"""
import argparse
import ctypes as C
import os
import sys

os.environ["B200ROMP_TC_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from romp_b200 import _lib
from romp_b200._lib import BF16
from romp_b200.graph import NetBuilder

NAMES = ["entry", "prologue", "pred_done", "weights", "first_tile", "last_mma", "epi_done", "exit"]


def build(cin, hw, B, n_ops, rs):
    nb = NetBuilder(0, "bf16", _lib.ENGINE_AUTO)
    ext = nb.tensor(hw, hw, cin, BF16, external=1)
    eye = np.zeros((cin, cin, 1, 1), np.float32)
    eye[np.arange(cin), np.arange(cin), 0, 0] = 1
    t = nb.conv(ext, eye, None)
    w = rs.normal(0, 1 / np.sqrt(cin * 9), (cin, cin, 3, 3)).astype(np.float32)
    b = rs.normal(0, 0.1, cin).astype(np.float32)
    for i in range(n_ops):
        t = nb.conv(t, w, b, relu=True, res=t if i % 2 else None)
    nb.finalize(B)
    return nb, ext


def timeit(nb, B, st, iters=20):
    for _ in range(3):
        nb.lib.b200romp_net_run(nb.net, B, C.c_void_p(st.cuda_stream))
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters):
        nb.lib.b200romp_net_run(nb.net, B, C.c_void_p(st.cuda_stream))
    e1.record(st)
    st.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cin", type=int, default=128)
    ap.add_argument("--hw", type=int, default=32)
    ap.add_argument("--ops", type=int, default=6)
    args = ap.parse_args()
    rs = np.random.RandomState(0)
    st = torch.cuda.Stream()
    print(f"## {args.cin}->{args.cin} 3x3 @{args.hw}^2, chain of {args.ops} convs in one CUDA graph")
    for B in (8, 16, 32, 64, 128):
        x = torch.randn(B, args.hw, args.hw, args.cin, device="cuda").bfloat16()
        nb0, e0 = build(args.cin, args.hw, B, 0, rs)
        nb1, e1 = build(args.cin, args.hw, B, args.ops, rs)
        for nb, e in ((nb0, e0), (nb1, e1)):
            nb.lib.b200romp_net_bind(nb.net, e, C.c_void_p(x.data_ptr()))
        t0, t1 = timeit(nb0, B, st), timeit(nb1, B, st)
        tiles = B * args.hw * args.hw // 128
        print(f"batch {B:4d}: {tiles:5d} tiles ({tiles / 148:5.1f} per SM)  {(t1 - t0) / args.ops:6.1f} us per op   {nb1.describe().splitlines()[1][-90:]}")
        if B == 64:
            n = args.ops + 1
            buf = np.zeros((n, 4, 16), np.uint64)
            _lib.check(nb1.lib.b200romp_net_read_stamps(nb1.net, buf.ctypes.data_as(C.c_void_p), n), "read_stamps")
            s = buf.astype(np.int64)
            t_ref = s[1, 0, 0]
            print("   timeline of the last replay, us relative to the entry of op 1's CTA 0 (CTA 0 = leader, CTA 1 = its peer, CTA 2 = next pair's leader)")
            print("   op cta " + " ".join(f"{k:>10s}" for k in NAMES))
            for op in range(1, n):
                for cta in (0, 1, 2):
                    row = " ".join(f"{(s[op, cta, k] - t_ref) / 1e3:10.2f}" if s[op, cta, k] else f"{'-':>10s}" for k in range(8))
                    print(f"   {op:2d} {cta:3d} {row}")


if __name__ == "__main__":
    main()
