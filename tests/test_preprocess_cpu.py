"""CPU: the host restatement of the device cubic resize (simple_hrnet_b200.preprocess) against cv2 -- the resize of the
reference's preprocessing, SimpleHRNet.py:216-220 / :356-360.  Bit-identical to OpenCV's own kernel
(cv2.setUseOptimized(False)); against the default cv2 call, which takes the vendor (IPP) path, at most one grey level
apart on a few per cent of the pixels."""
import cv2
import numpy as np
import pytest

from simple_hrnet_b200.preprocess import resize_cubic_reference

SHAPES = [(300, 210, 256, 192), (480, 640, 384, 288), (30, 20, 384, 288), (385, 289, 384, 288), (100, 37, 256, 192),
          (700, 500, 257, 191), (384, 288, 384, 288)]


@pytest.mark.parametrize("sh,sw,dh,dw", SHAPES)
def test_matches_opencv_own_kernel_bit_for_bit(sh, sw, dh, dw):
    img = np.random.default_rng(sh * 7 + sw).integers(0, 256, (2, sh, sw, 3), dtype=np.uint8)
    img[1, : sh // 2] = 255; img[1, sh // 2:] = 0          # saturation on both sides of an edge
    was = cv2.useOptimized()
    cv2.setUseOptimized(False)
    try:
        want = np.stack([cv2.resize(i, (dw, dh), interpolation=cv2.INTER_CUBIC) for i in img])
    finally:
        cv2.setUseOptimized(was)
    assert np.array_equal(resize_cubic_reference(img, dh, dw), want)


def test_distance_to_default_cv2_path_is_one_grey_level():
    img = np.random.default_rng(3).integers(0, 256, (1, 300, 210, 3), dtype=np.uint8)
    default = cv2.resize(img[0], (192, 256), interpolation=cv2.INTER_CUBIC)[None]
    d = np.abs(default.astype(np.int32) - resize_cubic_reference(img, 256, 192).astype(np.int32))
    assert d.max() <= 1 and (d != 0).mean() < 0.08
