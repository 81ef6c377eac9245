"""N>1 host logic on CPU: world_size-2 gloo processes exercise the frame partition and the single all-gather of
packed per-person records (romp_b200/shard.py); the gathered result must equal the unsharded one.  Covers the cases the
round-1 advisor flagged: one rank without persons next to a rank with persons (default AND non-default record width),
a rank without frames, nobody anywhere, and the overflow path of the rows hint - always with the same number of
collectives on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from romp_b200 import shard


def fake_outputs(frames, n_betas=10, n_verts=20, persons_of=lambda f: f % 3):
    """per-frame deterministic fake persons"""
    rows = [(local, f, k) for local, f in enumerate(frames) for k in range(persons_of(f))]
    if not rows:
        return None
    val = lambda f, k, w: (np.arange(w) + 1000 * f + 10 * k).astype(np.float32)
    return {"cam": np.stack([val(f, k, 3) for _, f, k in rows]),
            "smpl_thetas": np.stack([val(f, k, 72) for _, f, k in rows]),
            "smpl_betas": np.stack([val(f, k, n_betas) for _, f, k in rows]),
            "center_confs": np.stack([val(f, k, 1) for _, f, k in rows]),
            "cam_trans": np.stack([val(f, k, 3) for _, f, k in rows]),
            "joints": np.stack([val(f, k, 213).reshape(71, 3) for _, f, k in rows]),
            "pj2d_org": np.stack([val(f, k, 142).reshape(71, 2) for _, f, k in rows]),
            "verts": np.stack([val(f, k, 3 * n_verts).reshape(n_verts, 3) for _, f, k in rows]),
            "center_preds": np.stack([np.array([f, k], np.int64) for _, f, k in rows]),
            "pred_batch_ids": np.array([l for l, _, _ in rows], np.int64)}


PERSONS = {"mod3": lambda f: f % 3, "rank1_empty": lambda f: 2 if f < 2 else 0, "many": lambda f: 7}


def worker(rank, world, port, total, q, n_betas, persons, rows_hint):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_range(total, rank, world)
    layout = shard.romp_layout(True, n_betas, 20)
    out = fake_outputs(list(range(lo, hi)), n_betas, 20, PERSONS[persons])
    res = shard.all_gather_outputs(out, lo, world, layout=layout, rows_hint=rows_hint)
    # the pipelined form (persistent buffers, hint carried from step to step) must give the same answer, twice
    n = 0 if out is None else len(out["cam"])
    pipe = shard.ShardGather(world, layout, capacity=64, rows_hint=rows_hint)
    fields = ({k: torch.from_numpy(v) for k, v in out.items()} if out is not None else
              {name: torch.zeros((1,) + shp, dtype=dt) for name, shp, dt in layout.fields})
    for _ in range(2):
        res2 = pipe.result(pipe.submit(fields, n, lo), to_numpy=True)
        assert (res is None) == (res2 is None)
        if res is not None:
            assert all(np.array_equal(res[k], res2[k]) for k in res)
    q.put((rank, None if res is None else dict(res), pipe.collectives))
    dist.destroy_process_group()


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def run(world, total, n_betas=10, persons="mod3", rows_hint=64):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    ps = [ctx.Process(target=worker, args=(r, world, port, total, q, n_betas, persons, rows_hint)) for r in range(world)]
    [p.start() for p in ps]
    got = {}
    for _ in range(world):
        r, res, ncoll = q.get(timeout=120)
        got[r] = (res, ncoll)
    [p.join(timeout=60) for p in ps]
    assert len({nc for _, nc in got.values()}) == 1, "ranks issued different numbers of collectives"
    return {r: v[0] for r, v in got.items()}, got[0][1]


def check_equal(res, ref):
    assert np.array_equal(res["pred_batch_ids"], ref["pred_batch_ids"])          # global frame ids, ascending
    for k in ("cam", "smpl_thetas", "smpl_betas", "joints", "verts", "center_preds", "pj2d_org", "center_confs", "cam_trans"):
        assert np.array_equal(res[k], ref[k]), k
        assert res[k].dtype == ref[k].dtype
    assert res["body_pose"].shape[1] == 69 and res["global_orient"].shape[1] == 3


def test_shard_range_partitions():
    for total in (1, 7, 64, 512):
        for world in (1, 2, 3, 8):
            rs = [shard.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))


def test_record_layout_is_config_not_data():
    a, b = shard.romp_layout(True, 10), shard.romp_layout(True, 11)
    assert a.row_bytes % 16 == 0 and a.row_bytes >= (3 + 72 + 10 + 1 + 3 + 213 + 142 + 20670) * 4 + 24
    assert a != b and b.row_bytes >= a.row_bytes
    assert [n for n, _, _ in a.fields][-2:] == ["center_preds", "pred_batch_ids"]


def test_all_gather_equals_unsharded():
    total = 7
    got, ncoll = run(2, total)
    ref = fake_outputs(list(range(total)), 10, 20)
    ref["pred_batch_ids"] = np.array([f for f in range(total) for _ in range(f % 3)], np.int64)
    for r in (0, 1):
        check_equal(got[r], ref)
    assert ncoll == 2           # one per pipelined step, nothing else


def test_all_gather_with_an_empty_rank_and_nobody():
    got, _ = run(2, 1)            # rank 1 has no frame at all; frame 0 has 0 persons -> None everywhere
    assert got[0] is None and got[1] is None
    got, _ = run(2, 3)            # rank 0: frames 0,1 ; rank 1: frame 2
    assert got[0]["pred_batch_ids"].tolist() == [1, 2, 2]


def test_one_empty_and_one_full_rank_default_and_nondefault_width():
    """a rank that saw nobody must take exactly the collectives of a rank that saw persons, for the
    default record (10 betas) and for a non-default one (11 betas, the BEV width)."""
    for n_betas in (10, 11):
        got, ncoll = run(2, 4, n_betas=n_betas, persons="rank1_empty")      # rank 0: frames 0,1 (2 persons each); rank 1: none
        ref = fake_outputs([0, 1, 2, 3], n_betas, 20, PERSONS["rank1_empty"])
        ref["pred_batch_ids"] = np.array([0, 0, 1, 1], np.int64)
        for r in (0, 1):
            check_equal(got[r], ref)
        assert ncoll == 2


def test_rows_hint_overflow_regathers_collectively():
    got, ncoll = run(2, 4, persons="many", rows_hint=4)                      # 14 persons per rank > hint 4
    ref = fake_outputs([0, 1, 2, 3], 10, 20, PERSONS["many"])
    ref["pred_batch_ids"] = np.repeat(np.arange(4), 7).astype(np.int64)
    for r in (0, 1):
        check_equal(got[r], ref)
    assert ncoll == 3           # step 1: gather + exact re-gather; step 2 runs with the updated hint
