"""CPU oracle of row a1 (img_preprocess / padding_image, simple_romp/romp/utils.py:16-30).  TEST INFRASTRUCTURE ONLY.

``cv2.resize(..., INTER_CUBIC)`` on uint8 is the one third-party algorithm on this row.  This file restates OpenCV's OWN
8-bit bicubic resize (opencv/modules/imgproc/src/resize.cpp: interpolateCubic, the fixed-point coefficient tables of
cv::resize, HResizeCubic<uchar,int,short>, and the vectorised VResizeCubicVec_32s8u) in numpy, operation for operation.
Pinned by tests/test_preprocess_golden.py against tests/golden/preproc_opencv.npz = outputs of the REFERENCE's
img_preprocess with OpenCV's closed-source IPP fast path switched off (cv2.ipp.setUseIPP(False); written by
tests/golden/make_golden_preproc.py): bit-exact.  With IPP on (the default of pip wheels on x86) cv2's results differ from
OpenCV's own code by +-1 LSB on ~3 % of the pixels (CPU-dispatch dependent) - that fixture (preproc.npz) is matched to
+-1 LSB only.  OpenCV 4.13.0 here; the algorithm has been stable since 3.x.
"""
import numpy as np


def cubic_table(n_dst, n_src):
    """source start index and 11-bit fixed-point taps per destination coordinate (resize.cpp: cv::resize, interpolateCubic)."""
    scale = n_src / n_dst                                  # double, like scale_x = (double)ssize.width/dsize.width
    s0 = np.zeros(n_dst, np.int32)
    taps = np.zeros((n_dst, 4), np.int32)
    A, one = np.float32(-0.75), np.float32(1)
    for d in range(n_dst):
        fx = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(fx))
        x = np.float32(fx - np.float32(s))
        c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
        c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
        xm = one - x
        c2 = ((A + np.float32(2)) * xm - (A + np.float32(3))) * xm * xm + one
        c3 = one - c0 - c1 - c2
        taps[d] = np.rint(np.array([c0, c1, c2, c3], np.float32) * np.float32(2048)).astype(np.int32)   # saturate_cast<short>: half-even
        s0[d] = s
    return s0, taps


def resize_cubic_u8(img, size):
    """cv2.resize(img, (size, size), interpolation=cv2.INTER_CUBIC) for uint8 HxWxC, OpenCV's own code path."""
    H, W, _ = img.shape
    sx, ax = cubic_table(size, W)
    sy, ay = cubic_table(size, H)
    xi = np.clip(sx[:, None] + np.arange(-1, 3)[None], 0, W - 1)
    rows = (img[:, xi, :].astype(np.int64) * ax[None, :, :, None]).sum(2).astype(np.int32)      # HResizeCubic: int32, no shift
    yi = np.clip(sy[:, None] + np.arange(-1, 3)[None], 0, H - 1)
    S = rows[yi].astype(np.float32)                                                             # [dst_y, 4, dst_x, C]
    b = (ay.astype(np.float32) * (np.float32(1.0) / np.float32(2048 * 2048)))[:, :, None, None]
    t = S[:, 3] * b[:, 3]                                    # VResizeCubicVec_32s8u: fp32, separate mul and add, this order
    for k in (2, 1, 0):
        t = (S[:, k] * b[:, k]).astype(np.float32) + t
    return np.clip(np.rint(t), 0, 255).astype(np.uint8)


def padding_image(image):
    """utils.py:16-24."""
    h, w = image.shape[:2]
    side = max(h, w)
    pad = np.zeros((side, side, 3), dtype=np.uint8)
    top, left = int((side - h) // 2), int((side - w) // 2)
    pad[top:top + h, left:left + w] = image
    return pad, np.array([top, top + h, left, left + w, h, w], dtype=np.float32)


def img_preprocess(image_bgr, input_size=512):
    """utils.py:26-30 without OpenCV: BGR->RGB, centre zero pad, bicubic resize -> (uint8 [1,S,S,3], pad info)."""
    pad, info = padding_image(np.ascontiguousarray(image_bgr[:, :, ::-1]))
    return resize_cubic_u8(pad, input_size)[None], info
