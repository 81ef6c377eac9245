"""Multi-GPU check of the sharded path's single collective (romp_b200/shard.py ShardGather on NCCL: device-side pack kernel
+ one all_gather_into_tensor per step).  Run under torchrun on >= 2 GPUs:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_gather_check.py
Cases: every rank has persons; one rank has none (the empty rank must issue the same collectives); the rows
hint overflows (collectively decided re-gather); nobody anywhere.  Exit code 0 = all ranks agree with the expected result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romp_b200 import shard  # noqa: E402


def fields_for(layout, n, cap, seed, dev):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shp, dt in layout.fields:
        if dt == torch.float32:
            t = torch.randn((cap,) + shp, generator=g)
        else:
            t = torch.randint(0, 64, (cap,) + shp, generator=g, dtype=torch.int64)
        out[name] = t.to(dev)
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    layout = shard.romp_layout(True, 10)
    cap = 128
    B = 8
    ok = True
    for case, counts, hint in (("all", [5 + r for r in range(world)], 16), ("rank1_empty", [7 if r != 1 else 0 for r in range(world)], 16),
                               ("overflow", [40 + r for r in range(world)], 8), ("nobody", [0] * world, 16)):
        g = shard.ShardGather(world, layout, capacity=cap, rows_hint=hint)
        all_fields = [fields_for(layout, counts[r], cap, 100 + r, "cpu") for r in range(world)]
        mine = {k: v.to(dev) for k, v in all_fields[rank].items()}
        cnt = torch.tensor([counts[rank]], dtype=torch.int32, device=dev)
        for step in range(2):                       # second step runs with the hint learned from the first
            res = g.result(g.submit(mine, cnt, rank * B), to_numpy=True)
            if sum(counts) == 0:
                ok &= res is None
                continue
            ok &= res is not None
            for name, shp, dt in layout.fields:
                want = torch.cat([all_fields[r][name][:counts[r]] for r in range(world)]).numpy()
                if name == "pred_batch_ids":
                    want = want + np.concatenate([np.full(counts[r], r * B, np.int64) for r in range(world)])
                got = res[name]
                if got.shape != want.shape or not np.array_equal(got, want):
                    ok = False
                    print(f"[rank {rank}] case {case} step {step}: field {name} mismatch", flush=True)
        expected_collectives = {"all": 2, "rank1_empty": 2, "overflow": 3, "nobody": 2}[case]
        if g.collectives != expected_collectives:
            ok = False
            print(f"[rank {rank}] case {case}: {g.collectives} collectives, expected {expected_collectives}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("nccl_gather_check:", "OK" if int(flag) == 1 else "FAILED", f"(world {world})", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
