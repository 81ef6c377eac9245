"""Golden fixture for row f4 (temporal path): the REFERENCE's One-Euro smoothing (simple_romp/romp/utils.py:188-270:
create_OneEuroFilter, smooth_results, OneEuroFilter, LowPassFilter, smooth_global_rot_matrix) run on seeded sequences.

    python tests/golden/make_golden_one_euro.py          # build container only (needs /root/reference)

3 persons x 8 frames of (smpl_thetas[72], smpl_betas[10], cam[3]) random walks; each person has its own filter set with
smooth_coeff 3.0 (the default of --smooth_coeff).  Inputs and the reference's outputs are stored."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402

T, P = 8, 3


def sequences():
    rs = np.random.RandomState(21)
    thetas = np.cumsum(rs.normal(0, 0.08, size=(T, P, 72)), 0).astype(np.float32) + rs.normal(0, 0.5, size=(1, P, 72)).astype(np.float32)
    thetas[:, 1, :3] = np.array([3.0, 0.2, -0.1], np.float32) + thetas[:, 1, :3] * 0.1        # near-pi global rotation (branchy rotmat->aa)
    betas = np.cumsum(rs.normal(0, 0.05, size=(T, P, 10)), 0).astype(np.float32)
    cam = (np.array([0.8, 0.0, 0.1], np.float32) + np.cumsum(rs.normal(0, 0.02, size=(T, P, 3)), 0)).astype(np.float32)
    return thetas, betas, cam


def main():
    U = load_reference()["romp.utils"]
    thetas, betas, cam = sequences()
    filters = [U.create_OneEuroFilter(3.0) for _ in range(P)]
    o_t, o_b, o_c = np.zeros_like(thetas), np.zeros_like(betas), np.zeros_like(cam)
    for t in range(T):
        for p in range(P):
            a, b, c = U.smooth_results(filters[p], torch.from_numpy(thetas[t, p]), torch.from_numpy(betas[t, p]), torch.from_numpy(cam[t, p]))
            o_t[t, p], o_b[t, p], o_c[t, p] = a.numpy(), b.numpy(), c.numpy()
    np.savez_compressed(os.path.join(HERE, "one_euro.npz"), thetas=thetas, betas=betas, cam=cam, out_thetas=o_t, out_betas=o_b, out_cam=o_c)
    print("wrote one_euro.npz; max change by smoothing:", np.abs(o_t - thetas).max(), np.abs(o_c - cam).max())


if __name__ == "__main__":
    main()
