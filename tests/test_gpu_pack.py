"""Device-side record packing (csrc/pack.cu, b200romp_pack_rows) and the single-collective gather of the sharded path
(shard.ShardGather) on a real GPU: a one-rank NCCL group exercises pack kernel, header, all_gather_into_tensor, the pinned
header read-back and unpack with the person count read ON THE DEVICE.  (The multi-rank behaviour - empty ranks, overflow of
the rows hint - is covered by tests/test_shard_gloo.py on CPU and by tools/nccl_gather_check.py on 2+ GPUs.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

from romp_b200 import shard

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nccl_world1():
    if not dist.is_initialized():
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize("n,hint", [(37, 64), (0, 16), (90, 8)])
def test_pack_kernel_and_single_collective(nccl_world1, n, hint):
    layout = shard.romp_layout(True, 10)
    cap = 128
    g = torch.Generator().manual_seed(n)
    fields = {}
    for name, shp, dt in layout.fields:
        t = torch.randn((cap,) + shp, generator=g) if dt == torch.float32 else torch.randint(0, 512, (cap,) + shp, generator=g, dtype=torch.int64)
        fields[name] = t.cuda()
    count = torch.tensor([n], dtype=torch.int32, device="cuda")          # the kernel reads the count on the device
    sg = shard.ShardGather(1, layout, capacity=cap, rows_hint=hint)
    h = sg.submit(fields, count, frame_offset=640)
    counts, offsets = sg.counts(h)
    assert counts == [n] and offsets == [640]
    out = sg.result(h, to_numpy=True)
    if n == 0:
        assert out is None
        return
    assert sg.collectives == (2 if n > hint else 1)
    for name, shp, dt in layout.fields:
        want = fields[name][:n].cpu().numpy()
        if name == "pred_batch_ids":
            want = want + 640
        assert out[name].dtype == want.dtype and np.array_equal(out[name], want), name
    assert out["body_pose"].shape == (n, 69)
