#!/usr/bin/env python
"""bench.py - ROMP hot-path throughput on B200 (contract: see the task statement / DESIGN.md section 6).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # our arm  (frames/s, cfg2)
    python bench.py --impl reference --steps 2 --warmup 1          # reference arm: the CPU path on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...              # frames sharded over N ranks (weak scaling)

A step = one pass of the whole hot path (backbone+heads -> parse -> SMPL -> projection) over one batch of 64
synthetic 512x512 frames per GPU (BASELINE.json configs[1]).  Weights are seeded synthetic parameters with
the reference's schema; person detections are planted (1..10 per frame) because random weights detect nobody.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_FRAME = 85_710_602_240          # ROMP HRNet-32 + heads @512x512 (hooked on the reference)
SMPL_BYTES_PER_PERSON = 83_860
BATCH = 64


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16=d.get("bf16_tflops_sustained", 1472.5), bf16_burst=d.get("bf16_tflops", 1717.1),
                    hbm=d.get("hbm_gbs", 6484.3), src="measured")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def oracle_fps(sample_frames, threads):
    """The reference's CPU path (oracle port: fp32 torch-CPU restatement) on `sample_frames` frames."""
    import torch
    from oracle import romp_oracle as O
    from romp_b200 import synth
    torch.set_num_threads(threads)
    sd, pack = synth.romp_state_dict(0), synth.smpl_pack(0)
    frames = synth.synthetic_frames(sample_frames, seed=0)
    planted, _ = synth.plant_centers(sample_frames, seed=0)
    t0 = time.perf_counter()
    out = O.romp_forward(sd, pack, frames, center_override=planted)
    dt = time.perf_counter() - t0
    return sample_frames / dt, 0 if out is None else len(out["cam"])


def best_thread_count():
    """torch CPU convs do not scale to every core of a big host: time one frame at a few thread counts and keep
    the fastest ("all the host threads it can use")."""
    n = os.cpu_count() or 8
    cands = sorted({n, max(8, n // 2), max(8, n // 4), min(n, 32)}, reverse=True)
    oracle_fps(1, cands[0])                       # first-touch warm-up (weights generation, allocator)
    best = max(cands, key=lambda t: oracle_fps(1, t)[0])
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = best_thread_count()
    sample = 4
    for _ in range(args.warmup):
        oracle_fps(1, cores)
    t0 = time.perf_counter()
    persons = 0
    for _ in range(args.steps):
        _, n = oracle_fps(sample, cores)
        persons += n
    dt = time.perf_counter() - t0
    fps = args.steps * sample / dt
    line = {
        "impl": "reference", "metric": "frames/sec 512x512 ROMP-HRNet32 (whole hot path)", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2 ROMP HRNet-32 512x512, planted 1..10 persons/frame", "sample_frames_per_step": sample},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} frames/step of the cfg2 workload through oracle/romp_oracle.py (torch CPU fp32)"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "persons_per_step": persons / max(args.steps, 1),
    }
    emit(line)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from romp_b200 import ROMP, romp_settings, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    sd, pack = synth.romp_state_dict(0), synth.smpl_pack(0)
    s = romp_settings(["--GPU", str(local), "--precision", args.precision, "--max_batch", str(B)])
    model = ROMP(s, state_dict=sd, smpl_pack=pack)
    frames_host = torch.from_numpy(synth.synthetic_frames(B, seed=rank)).pin_memory()       # uint8 [B,512,512,3]
    planted_np, truth = synth.plant_centers(B, seed=rank)
    planted = torch.from_numpy(planted_np).cuda()
    persons = sum(len(t) for t in truth)
    frames_dev = frames_host.cuda()
    stream = model.stream
    offsets = [0, 512, 0, 512, 512, 512]

    def step_device():
        with torch.cuda.stream(stream):
            model.run_maps(frames_dev)
            model.run_post(B, offsets, planted)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(out):
        """the single collective of the sharded path: all-gather of the packed per-person outputs"""
        if world == 1:
            return out
        from romp_b200 import shard
        return shard.all_gather_outputs(out, rank * B, world)

    for _ in range(max(args.warmup, 3)):
        step_device()
        model.collect(to_numpy=False)
    # warm the streaming path too (pinned read-back mirrors of both slots, per-slot frame buffers, CUDA graphs)
    for _ in model.forward_batches((frames_host for _ in range(3)), center_override=planted):
        pass
    # ---------------- device-resident timing (value) + net-only timing (roofline) --------------------------
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    t_dev = e0.elapsed_time(e1) / 1e3
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record(stream)
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            model.run_maps(frames_dev)
    n1.record(stream)
    barrier()
    t_net = n0.elapsed_time(n1) / 1e3
    # ---------------- end to end through the public API with host buffers --------------------------------
    out = None
    if world > 1:
        from romp_b200 import shard
        pipe = shard.GatherPipeline(world, host_rank=0)     # warm: pinned mirrors of both slots, NCCL buffers
        for res in model.forward_batches((frames_host for _ in range(3)), center_override=planted, to_numpy=False):
            pipe.result(pipe.submit(res, rank * B))
    barrier()
    w0 = time.perf_counter()
    # public streaming API: per step H2D of that step's pinned frames, the whole path, D2H of the result dict;
    # copies of neighbouring steps overlap the kernels (forward_batches), results are consumed in order
    if world == 1:
        for res in model.forward_batches((frames_host for _ in range(args.steps)), center_override=planted, to_numpy=True):
            out = res
    else:
        # N ranks: each step's per-person outputs are all-gathered (NCCL) and read back on rank 0, pipelined one step deep
        pending = None
        for res in model.forward_batches((frames_host for _ in range(args.steps)), center_override=planted, to_numpy=False):
            h = pipe.submit(res, rank * B)
            if pending is not None:
                out = pipe.result(pending)
            pending = h
        out = pipe.result(pending)
    barrier()
    t_e2e = time.perf_counter() - w0
    clocks = sampler.finish() if rank == 0 else None
    d2h = 0 if out is None else int(sum(v.nbytes for v in out.values()))
    times = torch.tensor([t_dev, t_net, t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_net, t_e2e = times.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    fps = world * B * args.steps / t_dev
    net_tflops = B * args.steps * FLOP_PER_FRAME / t_net / 1e12          # per GPU
    nb, _ = model._net(2)
    desc = nb.describe()
    line = {
        "metric": "frames/sec 512x512 ROMP-HRNet32 (whole hot path)", "value": fps, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "cfg2 ROMP HRNet-32, batch 64 x 512x512 uint8 frames per GPU, planted 1..10 persons/frame",
                   "batch_per_gpu": B, "persons_per_gpu_step": persons, "parallelism": f"frames sharded x{world}",
                   "l2": "working set per step (activations > 1 GB) exceeds the 126 MB L2; no explicit flush"},
        "persons_per_sec": world * persons * args.steps / t_dev,
        "e2e": {"value": world * B * args.steps / t_e2e, "unit": "frames/s",
                "h2d_bytes_per_step": int(frames_host.numel()), "d2h_bytes_per_step": d2h},
        "gpu_launches": (nb.lib.b200romp_net_num_launches(nb.net) + 2 + 3 + 2) * args.steps,   # per timed device pass
        "roofline": {"bound": "tensor", "achieved": net_tflops, "peak": peaks["bf16"], "unit": "TFLOP/s",
                     "frac": net_tflops / peaks["bf16"], "traffic": None, "peak_source": peaks["src"] + " (sustained bf16)",
                     "kernel": "conv graph (backbone+heads), %d tcgen05 / %d simt ops" % (desc.count("tcgen05"), desc.count("simt   "))},
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        cores = best_thread_count()
        v, _ = oracle_fps(4, cores)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": cores, "kind": "port",
                                "sample": "4 frames of the cfg2 workload through oracle/romp_oracle.py (torch CPU fp32), after 1 warm-up frame"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_smpl(args):
    """BASELINE.json configs[4]: SMPL-only, 65,536 persons, GB/s against the HBM roofline (83,860 B/person)."""
    import ctypes as C
    import torch
    from romp_b200 import synth
    from romp_b200.main import SMPLParser
    torch.cuda.set_device(0)
    n = args.persons
    sm = SMPLParser(synth.smpl_pack(0), 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    betas = torch.randn(n, 10, generator=g).cuda()
    thetas = (torch.randn(n, 72, generator=g) * 0.3).cuda()
    verts = torch.empty(n, 6890, 3, device="cuda"); joints = torch.empty(n, 71, 3, device="cuda")
    ws = torch.empty(n, sm.ws_floats, device="cuda")
    st = torch.cuda.Stream()
    run = lambda: sm.forward(betas, thetas, n, None, False, ws, verts, joints, st.cuda_stream)
    for _ in range(max(args.warmup, 3)):
        run()
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(args.steps):
        run()
    e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    peaks = measured_peaks()
    gbs = n * SMPL_BYTES_PER_PERSON / ms / 1e6
    emit(({
        "metric": "persons/sec SMPL forward (verts + 71 joints)", "value": n / ms * 1e3, "unit": "persons/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg5 SMPL-only, %d persons, betas~N(0,1), thetas~N(0,0.3), synthetic SMPL pack" % n},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm"], "unit": "GB/s", "frac": gbs / peaks["hbm"],
                     "traffic": None, "peak_source": peaks["src"], "kernel": "smpl_pose + smpl_verts + smpl_joints"},
        "gpu_launches": 3 * args.steps}))


def run_bev(args):
    """BASELINE.json configs[2]: BEV HRNet-32 (+ bird's-eye-view head), batch 32 x 512x512, planted 3-D detections."""
    import torch
    from romp_b200 import synth
    from romp_b200.bev import BEV, bev_settings
    torch.cuda.set_device(0)
    B = args.batch if args.batch != BATCH else 32
    s = bev_settings(["--precision", args.precision, "--max_batch", str(B)])
    m = BEV(s, state_dict=synth.bev_state_dict(0), smpla_pack=synth.smpl_pack(0, num_betas=11), smil_pack=synth.smpl_pack(1))
    frames_host = torch.from_numpy(synth.synthetic_frames(B, seed=0)).pin_memory()
    frames_dev = frames_host.cuda()
    rs = np.random.RandomState(0)
    vol = rs.uniform(0, 0.05, size=(B, 64, 128, 128)).astype(np.float32)
    persons = 0
    for b in range(B):
        for _ in range(rs.randint(1, 11)):
            vol[b, rs.randint(0, 64), rs.randint(0, 128), rs.randint(0, 128)] = rs.uniform(0.3, 1.0)
            persons += 1
    vol = torch.from_numpy(vol).cuda()
    off = [0, 512, 0, 512, 512, 512]

    def step():
        with torch.cuda.stream(m.stream):
            m.run_model(frames_dev, vol)
            m.run_post(B, off)
    for _ in range(max(args.warmup, 3)):
        step()
        m.collect(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(m.stream)
    for _ in range(args.steps):
        step()
    e1.record(m.stream)
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / 1e3
    w0 = time.perf_counter()
    out = None
    for _ in range(args.steps):
        out = m.forward_batch(frames_host, center3d_override=vol)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - w0
    peaks = measured_peaks()
    fps = B * args.steps / t_dev
    emit(({
        "metric": "frames/sec 512x512 BEV-HRNet32 (whole hot path)", "value": fps, "unit": "frames/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": "cfg3 BEV HRNet-32 + BEV head, batch %d x 512x512 uint8, planted 1..10 persons/frame" % B,
                   "persons_planted": persons, "persons_out": 0 if out is None else int(len(out["cam"]))},
        "e2e": {"value": B * args.steps / t_e2e, "unit": "frames/s", "h2d_bytes_per_step": int(frames_host.numel()),
                "d2h_bytes_per_step": 0 if out is None else int(sum(v.nbytes for v in out.values()))},
        "roofline": {"bound": "tensor", "achieved": fps * 96_851_656_704 / 1e12, "peak": peaks["bf16"], "unit": "TFLOP/s",
                     "frac": fps * 96_851_656_704 / 1e12 / peaks["bf16"], "traffic": None, "peak_source": peaks["src"],
                     "kernel": "whole BEV step (conv graphs + BEV stages); 96.85 GFLOP/frame (SURVEY 8d)"}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", type=str, default="romp", choices=["romp", "bev", "smpl"],
                    help="romp = the contract's default (cfg2); bev = cfg3; smpl = cfg5 microbenchmark")
    ap.add_argument("--persons", type=int, default=65536)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=str, default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): library chatter written to file descriptor 1 while we run (NCCL prints
    # its version banner there on the first communicator) is diverted to stderr; emit() restores the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "smpl":
        run_smpl(args)
    elif args.workload == "bev":
        run_bev(args)
    else:
        run_ours(args)


_REAL_STDOUT = None


def emit(line):
    """print the result line on the process's original stdout"""
    text = json.dumps(line) + "\n"
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


if __name__ == "__main__":
    main()
