"""Exactly N whole-path steps (conv graph -> parse -> SMPL -> projection, cfg2 batch) bracketed by cudaProfilerStart/Stop, for
ncu launch lists:  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
                       --clock-control none --csv --log-file gpurun_out/launches.csv python tools/ncu_step.py --steps 1"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    from romp_b200 import ROMP, romp_settings, synth
    B = args.batch
    m = ROMP(romp_settings(["--precision", args.precision, "--max_batch", str(B)]), state_dict=synth.romp_state_dict(0),
             smpl_pack=synth.smpl_pack(0))
    frames = torch.from_numpy(synth.synthetic_frames(B, seed=0)).cuda()
    planted = torch.from_numpy(synth.plant_centers(B, seed=0)[0]).cuda()

    def step():
        with torch.cuda.stream(m.stream):
            m.run_maps(frames)
            m.run_post(B, [0, 512, 0, 512, 512, 512], planted)
        m.stream.synchronize()

    for _ in range(3):
        step()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(args.steps):
        step()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
