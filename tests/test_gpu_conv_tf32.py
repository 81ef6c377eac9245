"""TF32 tensor-core conv engine (tcgen05 kind::tf32 on fp32 NHWC tensors; csrc/conv_tc.cu, conv_tc_2cta.cu, conv_tc_s2.cu with
EB = 4) through the C ABI against a plain fp32 torch conv whose operands were rounded to TF32 exactly like the engine
does (cvt.rna.tf32.f32: 10 mantissa bits, ties away from zero).  Only the fp32 summation order differs, so the tolerance
is fp32-grade - an engine that let the tensor core TRUNCATE the activations (what kind::tf32 does with raw fp32 bits), or
that skipped the rounding of the weights, fails it by a factor of 10."""
import numpy as np
import pytest
import torch

from romp_b200 import _lib
from romp_b200._lib import F32
from tests.gpu_util import conv2d, conv_ref, round_tf32

pytestmark = pytest.mark.gpu

# name, k, cin, cout, H, W, B, relu, res, up, stride
CASES = [
    ("pair_k3_c32", 3, 32, 32, 32, 16, 2, 1, 0, 1, 1),
    ("pair_k3_c32_res_many_tiles", 3, 32, 32, 64, 64, 40, 1, 1, 1, 1),      # > 2 tiles per CTA and ring
    ("pair_k3_c32_n64", 3, 32, 64, 32, 32, 2, 0, 0, 1, 1),
    ("pair_k3_c64_res", 3, 64, 64, 32, 32, 3, 1, 1, 1, 1),
    ("pair_k3_c128_res", 3, 128, 128, 16, 16, 2, 1, 1, 1, 1),
    ("pair_k3_c256_res_many_tiles", 3, 256, 256, 16, 16, 24, 1, 1, 1, 1),
    ("pair_k3_c256_n32", 3, 256, 32, 32, 32, 1, 1, 0, 1, 1),
    ("single_k3_c64_odd_tiles", 3, 64, 64, 16, 8, 1, 0, 0, 1, 1),           # one tile per frame: no CTA pair possible
    ("k1_c64_n256_res", 1, 64, 256, 16, 16, 2, 1, 1, 1, 1),
    ("k1_c256_n64", 1, 256, 64, 32, 32, 3, 1, 0, 1, 1),
    ("k1_c128_n32_up4", 1, 128, 32, 16, 8, 2, 0, 0, 4, 1),
    ("k1_c32_n64", 1, 32, 64, 16, 16, 2, 0, 0, 1, 1),
    ("s2_c32_n32", 3, 32, 32, 64, 32, 2, 1, 0, 1, 2),
    ("s2_c32_n192", 3, 32, 192, 32, 32, 2, 1, 0, 1, 2),
    ("s2_c64_n128_res", 3, 64, 128, 64, 48, 2, 1, 1, 1, 2),
    ("s2_c128_n256_many_tiles", 3, 128, 256, 64, 64, 20, 1, 0, 1, 2),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tf32_conv_matches_torch_on_rounded_operands(case):
    name, k, cin, cout, H, W, B, relu, has_res, up, stride = case
    rs = np.random.RandomState(len(name) * 17 + cin)
    x = torch.from_numpy(rs.normal(0, 1, (B, H, W, cin)).astype(np.float32)).cuda()
    w = rs.normal(0, 1 / np.sqrt(cin * k * k), (cout, cin, k, k)).astype(np.float32)
    b = rs.normal(0, 0.5, cout).astype(np.float32)
    res = None
    if has_res:
        res = torch.from_numpy(rs.normal(0, 1, (B, H // stride * up, W // stride * up, cout)).astype(np.float32)).cuda()
    got = conv2d(x, w, b, stride=stride, relu=bool(relu), res=res, up=up, out_dtype=F32, engine=_lib.ENGINE_TF32).cpu()
    ref = conv_ref(round_tf32(x.cpu()), round_tf32(torch.from_numpy(w)).numpy(), b, stride=stride, relu=bool(relu), res=res, up=up)
    exact = conv_ref(x, w, b, stride=stride, relu=bool(relu), res=res, up=up)
    err, rounding = (got - ref).abs().max().item(), (exact - ref).abs().max().item()
    print(f"{name}: max|got - ref(tf32 operands)| {err:.2e}   (TF32 rounding itself moves the result by {rounding:.2e})")
    tol = 3e-5 + 1e-5 * ref.abs()
    bad = (got - ref).abs() > tol
    assert not bad.any(), f"{name}: {int(bad.sum())} of {bad.numel()} outputs off, max err {err:.3e}"
    assert rounding > 5 * err, "the case does not discriminate TF32 rounding from fp32"


def test_tf32_nchw_map_output_with_pow():
    """Head output convs in TF32 mode: [B,C,H,W] fp32 maps, 1.1**x on the cam-scale channel (main.py:112-113)."""
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.normal(0, 1, (2, 32, 32, 64)).astype(np.float32)).cuda()
    w = rs.normal(0, 0.125, (35, 64, 1, 1)).astype(np.float32)
    b = rs.normal(0, 0.5, 35).astype(np.float32)
    got = conv2d(x, w, b, out_dtype=F32, engine=_lib.ENGINE_TF32, out_nchw=1, pow_channel=0).cpu()
    ref = conv_ref(round_tf32(x.cpu()), round_tf32(torch.from_numpy(w)).numpy(), b, pow_channel=0).permute(0, 3, 1, 2)
    assert torch.allclose(got, ref, rtol=2e-5, atol=3e-5)


def test_tf32_engine_falls_back_to_simt_where_it_does_not_tile():
    """A shape outside the tensor-core tiling (Cin = 24) still computes - on the fp32 SIMT engine, i.e. without rounding."""
    rs = np.random.RandomState(2)
    x = torch.from_numpy(rs.normal(0, 1, (1, 16, 16, 24)).astype(np.float32)).cuda()
    w = rs.normal(0, 0.1, (32, 24, 3, 3)).astype(np.float32)
    got = conv2d(x, w, None, out_dtype=F32, engine=_lib.ENGINE_TF32).cpu()
    assert torch.allclose(got, conv_ref(x, w), rtol=1e-5, atol=1e-5)
