"""b200romp_net_add_sum (HRNet fuse-layer summation, simple_romp/romp/model.py:226-244) against plain torch fp32."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from romp_b200 import _lib
from romp_b200._lib import BF16, F32, SumDesc

pytestmark = pytest.mark.gpu
TD = {F32: torch.float32, BF16: torch.bfloat16}


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("H,W,Cc,ups", [(32, 32, 32, [2, 4, 8]), (16, 16, 64, [1, 2]), (8, 8, 256, [1]),
                                        (128, 128, 32, [2, 4, 8]), (32, 96, 32, [2])])   # last two: two chunks per thread / ragged second chunk
def test_fuse_sum_matches_torch(dtype, H, W, Cc, ups):
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    B = 3
    g = torch.Generator().manual_seed(1)
    base = torch.randn(B, H, W, Cc, generator=g).to(dev, TD[dtype])
    terms = [torch.randn(B, H // u, W // u, Cc, generator=g).to(dev, TD[dtype]) for u in ups]
    out = torch.empty(B, H, W, Cc, dtype=TD[dtype], device=dev)
    net = lib.b200romp_net_create(0)
    try:
        tb = lib.b200romp_net_add_tensor(net, H, W, Cc, dtype, 0, 1)
        tt = [lib.b200romp_net_add_tensor(net, H // u, W // u, Cc, dtype, 0, 1) for u in ups]
        to = lib.b200romp_net_add_tensor(net, H, W, Cc, dtype, 0, 1)
        d = SumDesc(to, tb, len(ups), (C.c_int * 4)(*(tt + [0] * (4 - len(tt)))), (C.c_int * 4)(*(ups + [1] * (4 - len(ups)))), 1)
        _lib.check(lib.b200romp_net_add_sum(net, C.byref(d)), "add_sum")
        _lib.check(lib.b200romp_net_finalize(net, B), "finalize")
        for t, x in zip([tb] + tt + [to], [base] + terms + [out]):
            _lib.check(lib.b200romp_net_bind(net, t, C.c_void_p(x.data_ptr())), "bind")
        _lib.check(lib.b200romp_net_run(net, B, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "run")
        torch.cuda.synchronize()
    finally:
        lib.b200romp_net_destroy(net)
    ref = base.float()
    for t, u in zip(terms, ups):          # same order as the kernel: base, term 0, 1, ...
        ref = ref + F.interpolate(t.float().permute(0, 3, 1, 2), scale_factor=u, mode="nearest").permute(0, 2, 3, 1)
    ref = F.relu(ref).to(TD[dtype])
    assert torch.equal(out, ref)          # fp32 adds in the same order, one final rounding: bit-exact


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_fuse_sum_term_channel_slices(dtype):
    """Terms that are channel slices of wider tensors (the merged 1x1 fuse convs of graph.py: one conv per source branch)."""
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    B, H, W, Cc = 2, 64, 64, 32
    g = torch.Generator().manual_seed(2)
    base = torch.randn(B, H, W, Cc, generator=g).to(dev, TD[dtype])
    wide = [(2, 128, 64), (4, 256, 96), (8, 64, 0)]        # up, channels of the term tensor, first channel of the slice
    terms = [torch.randn(B, H // u, W // u, ct, generator=g).to(dev, TD[dtype]) for u, ct, _ in wide]
    out = torch.empty(B, H, W, Cc, dtype=TD[dtype], device=dev)
    net = lib.b200romp_net_create(0)
    try:
        tb = lib.b200romp_net_add_tensor(net, H, W, Cc, dtype, 0, 1)
        tt = [lib.b200romp_net_add_tensor(net, H // u, W // u, ct, dtype, 0, 1) for u, ct, _ in wide]
        to = lib.b200romp_net_add_tensor(net, H, W, Cc, dtype, 0, 1)
        d = SumDesc(to, tb, 3, (C.c_int * 4)(*(tt + [0])), (C.c_int * 4)(2, 4, 8, 1), 1, (C.c_int * 4)(64, 96, 0, 0))
        _lib.check(lib.b200romp_net_add_sum(net, C.byref(d)), "add_sum")
        bad = SumDesc(to, tb, 1, (C.c_int * 4)(tt[2], 0, 0, 0), (C.c_int * 4)(8, 1, 1, 1), 1, (C.c_int * 4)(40, 0, 0, 0))
        assert lib.b200romp_net_add_sum(net, C.byref(bad)) < 0          # slice 40..72 leaves the 64-channel tensor
        _lib.check(lib.b200romp_net_finalize(net, B), "finalize")
        for t, x in zip([tb] + tt + [to], [base] + terms + [out]):
            _lib.check(lib.b200romp_net_bind(net, t, C.c_void_p(x.data_ptr())), "bind")
        _lib.check(lib.b200romp_net_run(net, B, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "run")
        torch.cuda.synchronize()
    finally:
        lib.b200romp_net_destroy(net)
    ref = base.float()
    for t, (u, _, co) in zip(terms, wide):
        ref = ref + F.interpolate(t[..., co:co + Cc].float().permute(0, 3, 1, 2), scale_factor=u, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(out, F.relu(ref).to(TD[dtype]))


def test_fuse_sum_rejects_bad_shapes():
    lib = _lib.load()
    net = lib.b200romp_net_create(0)
    try:
        a = lib.b200romp_net_add_tensor(net, 16, 16, 32, F32, 0, 1)
        b = lib.b200romp_net_add_tensor(net, 8, 8, 32, F32, 0, 1)
        o = lib.b200romp_net_add_tensor(net, 16, 16, 32, F32, 0, 1)
        d = SumDesc(o, a, 1, (C.c_int * 4)(b, 0, 0, 0), (C.c_int * 4)(4, 1, 1, 1), 1)   # 8*4 != 16
        assert lib.b200romp_net_add_sum(net, C.byref(d)) < 0
        assert b"term" in lib.b200romp_last_error()
    finally:
        lib.b200romp_net_destroy(net)
