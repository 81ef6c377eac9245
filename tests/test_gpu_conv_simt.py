"""SIMT conv engine (conv_simt.cu) through the C ABI against a plain fp32 torch conv of the same fused op."""
import numpy as np
import pytest
import torch

from romp_b200._lib import BF16, F32
from tests.gpu_util import conv2d, conv_ref

pytestmark = pytest.mark.gpu

CASES = [
    # cin, cout, k, stride, H, W, relu, res, up
    (64, 64, 3, 1, 16, 16, True, True, 1),
    (32, 32, 3, 1, 24, 40, True, False, 1),
    (256, 64, 1, 1, 16, 16, True, False, 1),
    (64, 32, 1, 1, 8, 8, False, True, 2),
    (256, 32, 1, 1, 8, 8, True, True, 8),
    (32, 64, 3, 2, 32, 32, False, True, 1),
    (128, 256, 3, 2, 16, 16, True, False, 1),
    (36, 20, 3, 1, 10, 12, False, False, 1),       # odd sizes / channel tails
]


@pytest.mark.parametrize("cin,cout,k,stride,H,W,relu,use_res,up", CASES)
def test_fp32_matches_torch(cin, cout, k, stride, H, W, relu, use_res, up):
    rs = np.random.RandomState(cin * 7 + cout + k + stride)
    x = torch.from_numpy(rs.normal(0, 1, (2, H, W, cin)).astype(np.float32)).cuda()
    w = rs.normal(0, 1 / np.sqrt(cin * k * k), (cout, cin, k, k)).astype(np.float32)
    b = rs.normal(0, 1, cout).astype(np.float32)
    Ho, Wo = ((H + 2 * (k // 2) - k) // stride + 1) * up, ((W + 2 * (k // 2) - k) // stride + 1) * up
    res = torch.from_numpy(rs.normal(0, 1, (2, Ho, Wo, cout)).astype(np.float32)).cuda() if use_res else None
    got = conv2d(x, w, b, stride=stride, relu=relu, res=res, up=up, out_dtype=F32).cpu()
    ref = conv_ref(x, w, b, stride=stride, relu=relu, res=res, up=up)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-5      # fp32 accumulate, different summation order only


def test_bf16_io_and_mixed_residual():
    rs = np.random.RandomState(3)
    x = torch.from_numpy(rs.normal(0, 1, (2, 16, 16, 64)).astype(np.float32)).cuda().bfloat16()
    w = torch.from_numpy(rs.normal(0, 1 / 24.0, (64, 64, 3, 3)).astype(np.float32)).bfloat16().float().numpy()
    res = torch.from_numpy(rs.normal(0, 1, (2, 16, 16, 64)).astype(np.float32)).cuda()
    got = conv2d(x, w, None, relu=True, res=res, out_dtype=BF16).float().cpu()
    ref = conv_ref(x, w, None, relu=True, res=res)
    assert (got - ref).abs().max().item() < 0.03       # one bf16 rounding of O(1..4) outputs


def test_u8_input_norm_nchw_and_pow():
    rs = np.random.RandomState(4)
    x = torch.from_numpy(rs.randint(0, 256, (2, 32, 32, 3)).astype(np.uint8)).cuda()
    w = rs.normal(0, 0.2, (7, 3, 3, 3)).astype(np.float32)
    b = rs.normal(0, 0.2, 7).astype(np.float32)
    got = conv2d(x, w, b, stride=2, input_norm=1, out_nchw=1, pow_channel=2).cpu()
    ref = conv_ref(x, w, b, stride=2, input_norm=1, pow_channel=2).permute(0, 3, 1, 2)
    assert (got - ref).abs().max().item() < 2e-5
