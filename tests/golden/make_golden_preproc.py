"""Golden fixture for row a1 (img_preprocess / padding_image, simple_romp/romp/utils.py:16-30) from the REFERENCE itself.

    python tests/golden/make_golden_preproc.py        # build container only (needs /root/reference)

Seeded random BGR images of several aspect ratios go through the reference's img_preprocess; the uint8-valued outputs
(small input_size keeps the fixture tiny; one case at the default 512 is stored as a checksum) and the pad-info vectors
are stored.  Only outputs are stored, nothing of the reference's source."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import load_reference  # noqa: E402

CASES = [(37, 53, 32), (120, 80, 32), (64, 64, 48), (5, 200, 32), (300, 400, 512)]
# second fixture (preproc_opencv.npz): the same reference function with OpenCV's IPP fast path switched off, i.e. OpenCV's
# own published resize algorithm - the bit-exactness target of the CUDA preprocessing kernel (csrc/preproc.cu)
CASES_OPENCV = CASES + [(511, 377, 512), (720, 1280, 512), (512, 512, 512), (90, 160, 64), (1080, 1920, 512)]


def images(cases=None):
    rs = np.random.RandomState(11)
    return [rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w, _ in (CASES if cases is None else cases)]


def checksum(x):
    xs = np.asarray(x).astype(np.float64)
    return np.array([xs.sum(), (xs * np.arange(xs.size).reshape(xs.shape) % 251).sum()])


def main():
    ref = load_reference()["romp.utils"]
    out = {}
    for i, (img, (h, w, size)) in enumerate(zip(images(), CASES)):
        x, pad = ref.img_preprocess(img, input_size=size)
        x = x.numpy()
        assert x.shape == (1, size, size, 3) and np.array_equal(x, np.round(x))      # float tensor holding uint8 values
        out[f"pad{i}"] = pad.numpy().astype(np.float32)
        if size <= 64:
            out[f"img{i}"] = x.astype(np.uint8)
        else:                                                                        # 512x512: checksum only
            out[f"sum{i}"] = np.array([x.astype(np.float64).sum(), (x.astype(np.float64) * np.arange(x.size).reshape(x.shape) % 251).sum()])
    np.savez_compressed(os.path.join(HERE, "preproc.npz"), **out)
    print("wrote preproc.npz", {k: v.shape for k, v in out.items()})
    # ---- OpenCV's own resize (IPP off)
    import cv2
    cv2.ipp.setUseIPP(False)
    out = {}
    for i, (img, (h, w, size)) in enumerate(zip(images(CASES_OPENCV), CASES_OPENCV)):
        x, pad = ref.img_preprocess(img, input_size=size)
        x = x.numpy()
        out[f"pad{i}"] = pad.numpy().astype(np.float32)
        if size <= 64:
            out[f"img{i}"] = x.astype(np.uint8)
        else:
            out[f"sum{i}"] = checksum(x)
            out[f"sample{i}"] = x.astype(np.uint8).reshape(-1)[::97].copy()       # every 97th byte, for diagnostics
    cv2.ipp.setUseIPP(True)
    np.savez_compressed(os.path.join(HERE, "preproc_opencv.npz"), **out)
    print("wrote preproc_opencv.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
