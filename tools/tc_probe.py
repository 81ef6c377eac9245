"""GPU probe for the tcgen05 conv engine: each case runs in its own subprocess (a trapped kernel poisons the
CUDA context), compares b200romp_conv2d(engine=TCGEN05) with a fp32 torch conv on bf16-rounded operands, and
on mismatch dumps arrays to gpurun_out/ for offline analysis.   python tools/tc_probe.py [--perf]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [
    # name, k, cin, cout, H, W, B, relu, res(0 none,1 f32,2 bf16), up, per_tap, out_bf16
    ("k1_c64_single", 1, 64, 64, 16, 8, 1, 0, 0, 1, 0, 0),
    ("k1_c64_multi", 1, 64, 64, 32, 32, 3, 1, 0, 1, 0, 1),
    ("k1_c256_n64", 1, 256, 64, 16, 16, 2, 0, 0, 1, 0, 0),
    ("k1_c128_n32_up4", 1, 128, 32, 16, 8, 2, 1, 1, 4, 0, 1),
    ("k1_c64_n256", 1, 64, 256, 16, 16, 2, 1, 2, 1, 0, 1),
    ("k3_c64_single", 3, 64, 64, 16, 8, 1, 0, 0, 1, 0, 0),
    ("k3_c64_single_pt", 3, 64, 64, 16, 8, 1, 0, 0, 1, 1, 0),
    ("k3_c64_multi", 3, 64, 64, 32, 24, 2, 1, 2, 1, 0, 1),
    ("k3_c64_multi_pt", 3, 64, 64, 32, 24, 2, 1, 2, 1, 1, 1),
    ("k3_c32", 3, 32, 32, 32, 16, 2, 1, 0, 1, 0, 1),
    ("k3_c32_pt", 3, 32, 32, 32, 16, 2, 1, 0, 1, 1, 1),
    ("k3_c128", 3, 128, 128, 16, 16, 2, 1, 2, 1, 0, 1),
    ("k3_c128_pt", 3, 128, 128, 16, 16, 2, 1, 2, 1, 1, 1),
    ("k3_c256", 3, 256, 256, 16, 16, 2, 1, 2, 1, 0, 1),
    ("k3_c256_n32", 3, 256, 32, 32, 32, 1, 1, 0, 1, 0, 1),
    # stride-2 (13th field): H, W are INPUT sizes
    ("s2_c64_single", 3, 64, 64, 32, 16, 1, 0, 0, 1, 0, 0, 2),
    ("s2_c64_multi", 3, 64, 128, 64, 48, 2, 1, 2, 1, 0, 1, 2),
    ("s2_c32_n32", 3, 32, 32, 64, 32, 2, 1, 1, 1, 0, 1, 2),
    ("s2_c32_n192", 3, 32, 192, 32, 32, 2, 1, 0, 1, 0, 1, 2),
    ("s2_c128_n256", 3, 128, 256, 32, 32, 2, 1, 2, 1, 0, 1, 2),
    ("s2_c256_n64", 3, 256, 64, 32, 32, 2, 1, 0, 1, 0, 1, 2),
]
PERF = [
    ("perf_k3_c64_64x64", 3, 64, 64, 64, 64, 64, 1, 2, 1, 0, 1),
    ("perf_k3_c32_128x128", 3, 32, 32, 128, 128, 64, 1, 2, 1, 0, 1),
    ("perf_k3_c128_32x32", 3, 128, 128, 32, 32, 64, 1, 2, 1, 0, 1),
    ("perf_k3_c256_16x16", 3, 256, 256, 16, 16, 64, 1, 2, 1, 0, 1),
    ("perf_k1_c64_n256_128", 1, 64, 256, 128, 128, 64, 1, 2, 1, 0, 1),
    ("perf_k3_c64_64x64_pt", 3, 64, 64, 64, 64, 64, 1, 2, 1, 1, 1),
    ("perf_k3_c64_128x128", 3, 64, 64, 128, 128, 64, 1, 2, 1, 0, 1),
    ("perf_k1_c64_n32_up2", 1, 64, 32, 64, 64, 64, 1, 1, 2, 0, 1),
    ("perf_k1_c256_n32_up8", 1, 256, 32, 16, 16, 64, 1, 1, 8, 0, 1),
    ("perf_s2_c64_128to64", 3, 64, 64, 128, 128, 64, 1, 0, 1, 0, 1, 2),
    ("perf_s2_c32_n64_128to64", 3, 32, 64, 128, 128, 64, 1, 0, 1, 0, 1, 2),
    ("perf_k1_c256_n64_128", 1, 256, 64, 128, 128, 64, 1, 0, 1, 0, 1),
    ("perf_s2_c256_n256_32to16", 3, 256, 256, 32, 32, 64, 1, 0, 1, 0, 1, 2),
    ("perf_s2_c128_n128_64to32", 3, 128, 128, 64, 64, 64, 1, 0, 1, 0, 1, 2),
]


def run_case(case, perf):
    import numpy as np
    import torch
    from romp_b200 import _lib
    from romp_b200._lib import BF16, F32
    from tests.gpu_util import conv2d, conv_ref
    name, k, cin, cout, H, W, B, relu, res_mode, up, per_tap, out_bf16 = case[:12]
    stride = case[12] if len(case) > 12 else 1
    os.environ["B200ROMP_TC_PER_TAP"] = "1" if per_tap else "0"
    rs = np.random.RandomState(len(name) * 131 + cin)
    x = torch.from_numpy(rs.normal(0, 1, (B, H, W, cin)).astype(np.float32)).cuda().bfloat16()
    w = torch.from_numpy(rs.normal(0, 1 / np.sqrt(cin * k * k), (cout, cin, k, k)).astype(np.float32)).bfloat16().float().numpy()
    b = rs.normal(0, 0.5, cout).astype(np.float32)
    res = None
    if res_mode:
        res = torch.from_numpy(rs.normal(0, 1, (B, H // stride * up, W // stride * up, cout)).astype(np.float32)).cuda()
        if res_mode == 2:
            res = res.bfloat16()
    od = BF16 if out_bf16 else F32
    got = conv2d(x, w, b, stride=stride, relu=bool(relu), res=res, up=up, out_dtype=od, engine=_lib.ENGINE_TCGEN05)
    info = {"case": name}
    if perf:
        # time with a prebuilt net (weights packed once).  The op under test reads an internal tensor produced by a
        # tcgen05 1x1 "feeder" conv; the feeder-only net is timed separately and subtracted.
        import ctypes as C
        from romp_b200.graph import NetBuilder

        def build(n_ops):
            nb = NetBuilder(0, "bf16", _lib.ENGINE_AUTO)
            ext = nb.tensor(H, W, cin, BF16, external=1)
            eye = np.zeros((cin, cin, 1, 1), np.float32); eye[np.arange(cin), np.arange(cin), 0, 0] = 1
            t = nb.conv(ext, eye, None)
            first = t
            for i in range(n_ops):
                chain = (cin == cout and stride == 1)
                t = nb.conv(t if chain else first, w, b, stride=stride, relu=bool(relu),
                            res=(t if (chain and res_mode) else None), up=up)
            nb.finalize(B)
            nb.lib.b200romp_net_bind(nb.net, ext, C.c_void_p(x.data_ptr()))
            return nb

        def timeit(nb, iters=10):
            st = torch.cuda.Stream()
            for _ in range(3):
                nb.lib.b200romp_net_run(nb.net, B, C.c_void_p(st.cuda_stream))
            st.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(iters):
                nb.lib.b200romp_net_run(nb.net, B, C.c_void_p(st.cuda_stream))
            e1.record(st)
            st.synchronize()
            return e0.elapsed_time(e1) / iters

        n_ops = 4
        nb0, nb1 = build(0), build(n_ops)
        t0, t1 = timeit(nb0), timeit(nb1)
        us = (t1 - t0) / n_ops * 1e3
        flops = 2.0 * B * (H // stride) * (W // stride) * cout * cin * k * k
        in_b = B * H * W * cin * 2
        out_b = B * (H // stride * up) * (W // stride * up) * cout * 2
        info.update(us_per_op=us, feeder_us=t0 * 1e3, tflops=flops / us / 1e6,
                    gbps_in_out=(in_b + out_b * (2 if res_mode else 1)) / us / 1e3, plan=nb1.describe().splitlines()[1][:170])
    ref = conv_ref(x.float(), w, b, stride=stride, relu=bool(relu), res=res, up=up)
    g = got.float().cpu()
    err = (g - ref).abs()
    tol = 0.02 + 0.02 * ref.abs() if out_bf16 else 2e-3 + 1e-3 * ref.abs()
    bad = err > tol
    info.update(max_err=float(err.max()), frac_bad=float(bad.float().mean()), ref_absmax=float(ref.abs().max()),
                ok=bool(bad.sum() == 0))
    if not info["ok"]:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        if g.numel() < 2_000_000:
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"tcfail_{name}.npz"), got=g.numpy(), ref=ref.numpy(),
                                x=x.float().cpu().numpy(), w=w)
        bb = bad.nonzero()
        info["bad_examples"] = bb[:8].tolist()
        info["bad_pix_in_tile"] = sorted(set(((int(r[1]) % 16) * 8 + int(r[2]) % 8) for r in bb[:4000].tolist()))[:40]
        info["bad_frames"] = sorted(set(int(r[0]) for r in bb[:4000].tolist()))
        info["bad_channels"] = sorted(set(int(r[3]) for r in bb[:4000].tolist()))[:40]
    print("PROBE " + json.dumps(info), flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        allc = CASES + PERF
        idx = int(sys.argv[2])
        run_case(allc[idx], idx >= len(CASES))
        return
    perf = "--perf" in sys.argv
    allc = CASES + (PERF if perf else [])
    for i, c in enumerate(allc):
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", str(i)], capture_output=True, text=True, timeout=120)
            lines = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            tail = (r.stderr or "")[-600:] + "\n".join(l for l in r.stdout.splitlines() if "b200romp" in l and "PROBE" not in l)[:400]
            print(lines[0] if lines else f"PROBE {{\"case\": \"{c[0]}\", \"ok\": false, \"rc\": {r.returncode}, \"err\": {json.dumps(tail)}}}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"PROBE {{\"case\": \"{c[0]}\", \"ok\": false, \"err\": \"timeout\"}}", flush=True)
        print(f"  ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
