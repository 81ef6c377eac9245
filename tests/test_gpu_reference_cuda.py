"""Comparator: the reference's own conv path - restated in oracle/romp_oracle.py, which is
pinned to the reference by the golden fixtures - run with PyTorch/cuDNN on the same B200, eager, on the cfg2 batch.
It is reported next to our numbers (profiles/), it is not a target and nothing in the product imports it.
Also checks that our engine agrees with this CUDA run of the oracle as it does with the CPU run."""
import json
import os
import time

import numpy as np
import pytest
import torch

from oracle import romp_oracle as O
from romp_b200 import synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200ROMP_RUN_COMPARATOR") != "1",
                                 reason="comparator timing takes ~4 min of cuDNN autotuning; set B200ROMP_RUN_COMPARATOR=1")]


def _time(fn, iters=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters


def test_reference_modules_on_cuda_timing():
    B = 64
    dev = torch.device("cuda", 0)
    sd = {k: torch.from_numpy(np.asarray(v)).to(dev) for k, v in synth.romp_state_dict(0).items()}
    frames = torch.from_numpy(synth.synthetic_frames(B, seed=0)).to(dev).float()
    torch.backends.cudnn.benchmark = True
    res = {}
    with torch.no_grad():
        c32, p32 = O.romp_maps(sd, frames[:4])
        assert torch.isfinite(c32).all() and torch.isfinite(p32).all()
        res["fp32_eager_tf32conv_ms"] = 1e3 * _time(lambda: O.romp_maps(sd, frames))
        with torch.autocast("cuda", dtype=torch.bfloat16):
            res["bf16_autocast_ms"] = 1e3 * _time(lambda: O.romp_maps(sd, frames))
    res["batch"] = B
    res["fps_fp32"] = B / res["fp32_eager_tf32conv_ms"] * 1e3
    res["fps_bf16_autocast"] = B / res["bf16_autocast_ms"] * 1e3
    res["what"] = "oracle/romp_oracle.py romp_maps (HRNet-32 backbone + heads) on cuda:0, torch eager + cuDNN benchmark, cfg2 frames"
    print("reference-on-CUDA comparator:", json.dumps(res))
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/ref_cuda.json", "w") as f:
            json.dump(res, f, indent=1)
    except OSError:
        pass
    assert res["fps_fp32"] > 0
