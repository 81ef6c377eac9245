"""N>1 host logic on CPU: world_size-2 gloo processes exercise the frame partition and the single all-gather of
packed outputs (romp_b200/shard.py); the gathered result must equal the unsharded one."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from romp_b200 import shard


def fake_outputs(frames, seed):
    """per-frame deterministic fake persons (frame f has f % 3 persons)"""
    rs = np.random.RandomState(seed)
    rows = []
    for local, f in enumerate(frames):
        for k in range(f % 3):
            rows.append((local, f, k))
    if not rows:
        return None
    n = len(rows)
    val = lambda f, k, w: (np.arange(w) + 1000 * f + 10 * k).astype(np.float32)
    out = {"cam": np.stack([val(f, k, 3) for _, f, k in rows]),
           "smpl_thetas": np.stack([val(f, k, 72) for _, f, k in rows]),
           "smpl_betas": np.stack([val(f, k, 10) for _, f, k in rows]),
           "center_confs": np.stack([val(f, k, 1) for _, f, k in rows]),
           "cam_trans": np.stack([val(f, k, 3) for _, f, k in rows]),
           "joints": np.stack([val(f, k, 213).reshape(71, 3) for _, f, k in rows]),
           "pj2d_org": np.stack([val(f, k, 142).reshape(71, 2) for _, f, k in rows]),
           "verts": np.stack([val(f, k, 60).reshape(20, 3) for _, f, k in rows]),
           "center_preds": np.stack([np.array([f, k], np.int64) for _, f, k in rows]),
           "pred_batch_ids": np.array([l for l, _, _ in rows], np.int64)}
    return out


def worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(total, rank, world)
    out = fake_outputs(list(range(lo, hi)), rank)
    res = shard.all_gather_outputs(out, lo, world)
    pipe = shard.GatherPipeline(world)                 # the pipelined form must give the same answer (sync on gloo)
    res2 = pipe.result(pipe.submit(out, lo))
    assert (res is None) == (res2 is None)
    if res is not None:
        assert all(np.array_equal(res[k], res2[k]) for k in res)
    q.put((rank, None if res is None else {k: v for k, v in res.items()}))
    dist.destroy_process_group()


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def run(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    ps = [ctx.Process(target=worker, args=(r, world, port, total, q)) for r in range(world)]
    [p.start() for p in ps]
    got = dict(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in ps]
    return got


def test_shard_range_partitions():
    for total in (1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            rs = [shard.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))


def test_all_gather_equals_unsharded():
    total = 7
    got = run(2, total)
    ref = fake_outputs(list(range(total)), 0)
    for r in (0, 1):
        res = got[r]
        assert np.array_equal(res["pred_batch_ids"], ref["pred_batch_ids"])      # global frame ids, ascending
        for k in ("cam", "smpl_thetas", "joints", "verts", "center_preds", "pj2d_org"):
            assert np.array_equal(res[k], ref[k]), k
        assert res["body_pose"].shape[1] == 69


def test_all_gather_with_an_empty_rank_and_nobody():
    got = run(2, 1)            # rank 1 has no frame at all; frame 0 has 0 persons -> None everywhere
    assert got[0] is None and got[1] is None
    got = run(2, 3)            # rank 0: frames 0,1 ; rank 1: frame 2
    assert got[0]["pred_batch_ids"].tolist() == [1, 2, 2]
