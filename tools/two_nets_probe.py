"""Does running the conv graph as two independent half-batch nets on two streams fill the layer-boundary bubbles?
(the fixed cost per layer is a drain -> visibility -> first-load chain during which the SMs idle.)
Times, with CUDA events:  one net x 64 frames   vs   two nets x 32 frames replayed concurrently on two streams."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from romp_b200 import ROMP, romp_settings, synth


def main():
    B = 64
    sd, smpl = synth.romp_state_dict(0), synth.smpl_pack(0)
    full = ROMP(romp_settings(["--precision", "bf16", "--max_batch", str(B)]), state_dict=sd, smpl_pack=smpl)
    halves = [ROMP(romp_settings(["--precision", "bf16", "--max_batch", str(B // 2)]), state_dict=sd, smpl_pack=smpl) for _ in range(2)]
    frames = torch.from_numpy(synth.synthetic_frames(B, seed=0)).cuda()
    fh = [frames[:B // 2].contiguous(), frames[B // 2:].contiguous()]
    main_s = torch.cuda.Stream()

    def run_full():
        with torch.cuda.stream(full.stream):
            full.run_maps(frames)

    def run_halves():
        for m, f in zip(halves, fh):
            with torch.cuda.stream(m.stream):
                m.run_maps(f)

    def timeit(fn, streams, iters=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for s in streams:
            s.wait_stream(main_s)
        for _ in range(iters):
            fn()
        for s in streams:
            main_s.wait_stream(s)
        e1.record(main_s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    t_full = timeit(run_full, [full.stream])
    t_half = timeit(run_halves, [m.stream for m in halves])
    def run_one():
        with torch.cuda.stream(halves[0].stream):
            halves[0].run_maps(fh[0])
    t_one_half = timeit(run_one, [halves[0].stream])
    print(f"one net x {B}: {t_full:.3f} ms/step ({B / t_full * 1e3:.0f} frames/s)")
    print(f"two nets x {B // 2} on two streams: {t_half:.3f} ms/step ({B / t_half * 1e3:.0f} frames/s)")
    print(f"one net x {B // 2} alone: {t_one_half:.3f} ms/step ({B / 2 / t_one_half * 1e3:.0f} frames/s)")
    c0, p0 = full.shared["center_maps"][:B].clone(), None
    torch.cuda.synchronize()
    c1 = torch.cat([m.shared["center_maps"][:B // 2] for m in halves])
    print("max |center diff| full vs halves:", float((c0 - c1).abs().max()))


if __name__ == "__main__":
    main()
