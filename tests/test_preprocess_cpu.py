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


def test_pil_bilinear_restatement_equals_pillow():
    """`ToPILImage -> Resize((H, W))` of the multi-person crop transform (SimpleHRNet.py:166-171): the host restatement of
    Pillow's antialiased bilinear resample (coefficient tables + the two rounded passes) that the device kernel evaluates
    must equal Pillow -- and torchvision's transform -- bit for bit, for up- and down-scaling crops."""
    from PIL import Image
    from torchvision.transforms import transforms
    from simple_hrnet_b200.preprocess import pil_bilinear_tables, resize_pil_bilinear_reference
    rng = np.random.default_rng(5)
    for t, (h, w) in enumerate([(23, 17), (96, 72), (384, 288), (385, 287), (700, 260), (150, 480), (511, 333), (64, 1000)]):
        oh, ow = (384, 288) if t % 2 == 0 else (256, 192)
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(resize_pil_bilinear_reference(img, oh, ow), ref), (h, w, oh, ow)
    tf = transforms.Compose([transforms.ToPILImage(), transforms.Resize((384, 288))])
    img = rng.integers(0, 256, (411, 233, 3), dtype=np.uint8)
    assert np.array_equal(resize_pil_bilinear_reference(img, 384, 288), np.asarray(tf(img)))
    b, k = pil_bilinear_tables(288, 288)                      # same size: identity taps
    assert np.array_equal(b[:, 0], np.arange(288)) and np.all(k[:, 0] == 1 << 22) and np.all(k[:, 1:] == 0)
    b, k = pil_bilinear_tables(1000, 288)                     # coefficients of every output sum to 2^22 up to rounding
    assert np.all(np.abs(k.sum(1) - (1 << 22)) <= k.shape[1])
