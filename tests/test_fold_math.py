"""The algebra behind two graph-level rewrites of the conv path, checked on the CPU with plain torch convs (the CUDA
implementations are covered by tests/test_gpu_conv_tc.py and tests/test_gpu_e2e.py):

* pixel-pair folding (csrc/net.cu: fold_pixel_pairs): a 3x3 stride-1 conv C->C on [B,H,W,C] equals a 3x3 conv 2C->2C on the
  byte-identical view [B,H,W/2,2C] with W2[dx*C+co][h*C+ci][ky][s+1] = W[co][ci][ky][2s+h-dx+1] (zero when that kx is outside
  0..2), including the zero padding at the left / right borders;
* the layer1.0 skip path (graph.py: build_backbone): conv3(t) + downsample(x) = one 1x1 conv over the channel concatenation
  [t | x] with weights [W3 | Wd] and bias b3 + bd (reference: romp/lib/models/resnet_50.py-style Bottleneck,
  simple_romp/romp/model.py:93-120)."""
import numpy as np
import torch
import torch.nn.functional as F


def fold_pixel_pairs(w, b):
    co_n, ci_n = w.shape[0], w.shape[1]
    w2 = np.zeros((2 * co_n, 2 * ci_n, 3, 3), np.float32)
    for dx in range(2):
        for h in range(2):
            for s in (-1, 0, 1):
                kx = 2 * s + h - dx + 1
                if 0 <= kx <= 2:
                    w2[dx * co_n:(dx + 1) * co_n, h * ci_n:(h + 1) * ci_n, :, s + 1] = w[:, :, :, kx]
    return w2, np.concatenate([b, b])


def test_pixel_pair_folding_equals_the_conv():
    rs = np.random.RandomState(0)
    B, H, W, Cc = 2, 16, 32, 32
    x = torch.from_numpy(rs.normal(0, 1, (B, H, W, Cc)).astype(np.float32))
    w = rs.normal(0, 0.1, (Cc, Cc, 3, 3)).astype(np.float32)
    b = rs.normal(0, 0.5, Cc).astype(np.float32)
    ref = F.conv2d(x.permute(0, 3, 1, 2), torch.from_numpy(w), torch.from_numpy(b), padding=1).permute(0, 2, 3, 1)
    w2, b2 = fold_pixel_pairs(w, b)
    xv = x.reshape(B, H, W // 2, 2 * Cc)                                   # same bytes: pixel pairs as 64-channel pixels
    got = F.conv2d(xv.permute(0, 3, 1, 2), torch.from_numpy(w2), torch.from_numpy(b2), padding=1).permute(0, 2, 3, 1)
    assert torch.allclose(got.reshape(B, H, W, Cc), ref, atol=2e-5)
    # the blocks the kernel skips (kmask) are exactly the all-zero ones: s = -1 uses only the second pixel, s = +1 only the first
    assert not w2[:, :Cc, :, 0].any() and not w2[:, Cc:, :, 2].any()
    assert w2[:, Cc:, :, 0].any() and w2[:, :Cc, :, 2].any() and w2[:, :, :, 1].any()


def test_skip_concat_equals_conv3_plus_downsample():
    rs = np.random.RandomState(1)
    t = torch.from_numpy(rs.normal(0, 1, (2, 64, 8, 8)).astype(np.float32))
    x = torch.from_numpy(rs.normal(0, 1, (2, 64, 8, 8)).astype(np.float32))
    w3, wd = (torch.from_numpy(rs.normal(0, 0.1, (256, 64, 1, 1)).astype(np.float32)) for _ in range(2))
    b3, bd = (torch.from_numpy(rs.normal(0, 0.5, 256).astype(np.float32)) for _ in range(2))
    ref = F.relu(F.conv2d(t, w3, b3) + F.conv2d(x, wd, bd))
    got = F.relu(F.conv2d(torch.cat([t, x], 1), torch.cat([w3, wd], 1), b3 + bd))
    assert torch.allclose(got, ref, atol=2e-5)
