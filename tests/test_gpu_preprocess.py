"""Row a1/f2 on the GPU: b200romp_preprocess_bgr (csrc/preproc.cu: BGR->RGB + centre zero pad + bicubic resize in one
kernel) through the C ABI - bit-exact against the reference's img_preprocess run on OpenCV's own resize
(tests/golden/preproc_opencv.npz), bit-exact against the numpy oracle on further odd shapes, within +-1 LSB of the
IPP-accelerated fixture; and ROMP.forward(image_bgr) - which never touches OpenCV - equals the batched entry point fed
with the same preprocessed frame."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from oracle import preproc_oracle as P
from romp_b200 import ROMP, _lib, romp_settings, synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_preproc import CASES, CASES_OPENCV, checksum, images  # noqa: E402

pytestmark = pytest.mark.gpu


def gpu_preprocess(img, size):
    lib = _lib.load()
    d = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    out = torch.empty((size, size, 3), dtype=torch.uint8, device="cuda")
    pad = (C.c_float * 6)()
    _lib.check(lib.b200romp_preprocess_bgr(C.c_void_p(d.data_ptr()), img.shape[0], img.shape[1], 3 * img.shape[1], size,
                                           C.c_void_p(out.data_ptr()), pad, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out.cpu().numpy()[None], np.array(list(pad), np.float32)


def test_kernel_is_bit_exact_with_reference_on_opencv_own_resize():
    g = np.load(os.path.join(HERE, "golden", "preproc_opencv.npz"))
    g_ipp = np.load(os.path.join(HERE, "golden", "preproc.npz"))
    for i, (img, (h, w, size)) in enumerate(zip(images(CASES_OPENCV), CASES_OPENCV)):
        x, pad = gpu_preprocess(img, size)
        assert np.array_equal(pad, g[f"pad{i}"]), f"pad info case {i}"
        if f"img{i}" in g:
            assert np.array_equal(x, g[f"img{i}"]), f"case {i} ({h}x{w} -> {size})"
        else:
            assert np.array_equal(x.reshape(-1)[::97], g[f"sample{i}"]), f"case {i} ({h}x{w} -> {size})"
            assert np.array_equal(checksum(x), g[f"sum{i}"]), f"case {i}"
        if i < len(CASES) and f"img{i}" in g_ipp.files:
            d = np.abs(x.astype(int) - g_ipp[f"img{i}"].astype(int))
            assert d.max() <= 1 and (d != 0).mean() < 0.15


@pytest.mark.parametrize("h,w,size", [(33, 97, 64), (97, 33, 64), (1, 1, 32), (2, 301, 96), (257, 255, 128), (640, 480, 512)])
def test_kernel_equals_numpy_oracle(h, w, size):
    img = np.random.RandomState(h * 1000 + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
    x, pad = gpu_preprocess(img, size)
    xo, po = P.img_preprocess(img, size)
    assert np.array_equal(pad, po) and np.array_equal(x, xo)


def test_forward_uses_the_gpu_preprocess_and_equals_forward_batch(monkeypatch):
    sd, pack = synth.romp_state_dict(0), synth.smpl_pack(0)
    img = np.random.RandomState(3).randint(0, 256, (300, 400, 3)).astype(np.uint8)
    frames = P.img_preprocess(img, 512)[0]
    c, _ = __import__("oracle.romp_oracle", fromlist=["x"]).romp_maps(sd, frames)
    sd2, _, _ = synth.calibrate_center_head(sd, c.numpy(), max_per_frame=6)
    m = ROMP(romp_settings(["--precision", "fp32", "--max_batch", "1"]), state_dict=sd2, smpl_pack=pack)
    import romp_b200.main as M
    monkeypatch.setattr(M, "img_preprocess", lambda *a, **k: (_ for _ in ()).throw(AssertionError("forward() must not call the OpenCV path")))
    fd, pad = m.preprocess(img)
    m.stream.synchronize()
    assert np.array_equal(fd.cpu().numpy()[None], frames) and pad.tolist() == [50, 350, 0, 400, 300, 400]
    out = m(img)
    ref = m.forward_batch(torch.from_numpy(frames), offsets=pad)
    assert out is not None and ref is not None and len(out["cam"]) >= 1
    ref.pop("pred_batch_ids")
    assert set(out) == set(ref)
    for k in out:
        assert np.array_equal(out[k], ref[k]), k
