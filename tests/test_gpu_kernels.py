"""`-m gpu` parity tests of the individual CUDA kernels, called through the C ABI."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import hrnet_oracle as O
from simple_hrnet_b200 import _lib
from tests import gpu_util as G

pytestmark = pytest.mark.gpu

# (n, ih, iw, cin, cout, k, stride): covers kc = 64 / 32 / 16 (swizzle 128/64/32), both strides, 1x1 and 3x3,
# N-split (cout 384, 512), M not a multiple of 128, tiles straddling rows and images, tiny maps (12x9, 8x6).
CONV_SHAPES = [
    (2, 16, 16, 64, 64, 1, 1),      # plain GEMM, one k-block
    (2, 16, 16, 64, 64, 3, 1),      # 3x3 im2col, SW128
    (3, 24, 18, 192, 192, 3, 1),    # W48 branch 2
    (5, 12, 9, 384, 384, 3, 1),     # W48 branch 3: odd map, N split 2x192, M = 540 (ragged last tile)
    (2, 48, 36, 96, 96, 3, 1),      # W48 branch 1: kc = 32 (SW64)
    (1, 96, 72, 48, 48, 3, 1),      # W48 branch 0: kc = 16 (SW32)
    (2, 64, 48, 32, 32, 3, 1),      # W32 branch 0: kc = 32
    (3, 8, 6, 256, 256, 3, 1),      # W32 branch 3: 8x6 map
    (2, 96, 72, 48, 96, 3, 2),      # stride-2 fuse/transition conv
    (2, 24, 18, 192, 384, 3, 2),    # stride-2 into the 12x9 map
    (2, 32, 32, 64, 64, 3, 2),      # stem conv2 shape class
    (2, 24, 18, 192, 48, 1, 1),     # 1x1 up-path conv
    (2, 24, 24, 256, 64, 1, 1),     # bottleneck 1x1 reduce
    (2, 24, 24, 64, 256, 1, 1),     # bottleneck 1x1 expand
    (1, 16, 12, 256, 512, 1, 2),    # PoseResNet 1x1 stride-2 downsample, N split
    (1, 8, 8, 16, 16, 3, 1),        # smallest legal channels; M = 64 < one tile
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_matches_cpu(shape):
    case = G.conv_case(*shape, relu=True, residual=(shape[6] == 1 and shape[3] == shape[4]), out_f32=False, seed=1)
    out = G.run_conv(case, use_tc=True)
    err = (out - case["ref"]).abs().max().item()
    assert err <= G.conv_tolerance(case), f"max abs err {err}"


# halo-patch kernel (use_tc=2): 3x3 s1 with resident weights; chunk mixes 64 / 32+16 / 64+32 / 4x64, ragged tile
# columns (W=36 -> 5 tiles of 8), ragged rows (H=40 -> 3 tiles of 16), single-tile and multi-image cases.
PATCH_SHAPES = [
    (1, 96, 72, 48, 48, 3, 1),      # W48 branch 0 (chunks 32 + 16)
    (2, 48, 36, 96, 96, 3, 1),      # W48 branch 1 (chunks 64 + 32), ragged tile columns
    (2, 96, 72, 64, 64, 3, 1),      # layer1 3x3
    (2, 64, 48, 32, 32, 3, 1),      # W32 branch 0
    (3, 32, 24, 64, 64, 3, 1),      # W32 branch 1
    (2, 64, 48, 256, 32, 3, 1),     # W32 transition1.0 (4 chunks of 64)
    (2, 96, 72, 256, 48, 3, 1),     # W48 transition1.0: weights (221 KB) streamed with the patch slots
    (2, 32, 16, 192, 64, 3, 1),     # streamed weights, 3 chunks
    (1, 16, 8, 16, 16, 3, 1),       # exactly one tile
    (5, 48, 24, 48, 96, 3, 1),      # Cin != Cout, 3x3 tiles per image
]


@pytest.mark.parametrize("shape", PATCH_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_patch_matches_cpu(shape):
    case = G.conv_case(*shape, relu=True, residual=(shape[3] == shape[4]), out_f32=False, seed=4)
    out = G.run_conv(case, use_tc=2)
    err = (out - case["ref"]).abs().max().item()
    assert err <= G.conv_tolerance(case), f"max abs err {err}"
    assert not torch.isnan(out).any()


def test_conv_patch_equals_im2col_kernel_bitwise_f32():
    """Same K order is not guaranteed, but with fp32 output and no epilogue both tcgen05 kernels must agree to
    accumulation-order noise, and the patch kernel must reject shapes it does not cover."""
    case = G.conv_case(2, 48, 36, 96, 96, 3, 1, relu=False, residual=False, out_f32=True, seed=6)
    a, b = G.run_conv(case, use_tc=1), G.run_conv(case, use_tc=2)
    assert (a - b).abs().max().item() < 1e-4
    bad = G.conv_case(1, 24, 18, 192, 192, 3, 1, relu=False, residual=False, out_f32=True, seed=6)
    with pytest.raises(_lib.HrnetError, match="not eligible"):
        G.run_conv(bad, use_tc=2)


@pytest.mark.parametrize("shape", [(3, 24, 18, 192, 192, 3, 1), (5, 12, 9, 384, 384, 3, 1), (2, 24, 18, 192, 384, 3, 2),
                                   (1, 16, 12, 256, 512, 1, 2), (2, 48, 36, 96, 96, 3, 1)], ids=lambda s: "x".join(map(str, s)))
def test_conv_tc_pair_mode_matches_cpu(shape):
    """CTA-pair mode (tcgen05 cta_group::2, opt-in via HRNET_TUNE_IGEMM_PAIR=2): M=256 MMAs over a 2-CTA cluster, half
    weight tile per CTA, odd tile counts (ghost CTA), N split."""
    case = G.conv_case(*shape, relu=True, residual=(shape[6] == 1 and shape[3] == shape[4]), out_f32=False, seed=8)
    _lib.set_debug_tune({_lib.TUNE_IGEMM_PAIR: 2})
    try:
        out = G.run_conv(case, use_tc=1)
    finally:
        _lib.set_debug_tune(None)
    err = (out - case["ref"]).abs().max().item()
    assert err <= G.conv_tolerance(case), f"max abs err {err}"


@pytest.mark.parametrize("shape", [(2, 96, 72, 48, 48), (3, 48, 36, 96, 96), (1, 16, 8, 64, 64), (5, 48, 24, 48, 96),
                                   (1, 32, 16, 128, 128)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("epi", ["direct", "tma", "coal"])
def test_conv_patch_pair_mode_matches_cpu(shape, epi):
    """Halo-patch kernel in CTA-pair mode (opt-in via HRNET_TUNE_PATCH_PAIR_MIN_COUT): M=256 MMAs over two tiles, half of
    the resident weight rows per CTA, odd tile counts (ghost CTA: (1,16,8) has one tile, (5,48,24) has 45), with the
    direct, the TMA-store and the warp-staged coalesced epilogue."""
    n, h, w, cin, cout = shape
    case = G.conv_case(n, h, w, cin, cout, 3, 1, relu=True, residual=(cin == cout), out_f32=False, seed=9)
    _lib.set_debug_tune({_lib.TUNE_PATCH_PAIR_MIN_COUT: 16,
                         _lib.TUNE_EPILOGUE: {"direct": _lib.EPI_DIRECT, "tma": _lib.EPI_TMA, "coal": _lib.EPI_COAL}[epi]})
    try:
        out = G.run_conv(case, use_tc=2)
    finally:
        _lib.set_debug_tune(None)
    err = (out - case["ref"]).abs().max().item()
    assert err <= G.conv_tolerance(case), f"max abs err {err}"


@pytest.mark.parametrize("epi", ["coal", "batch"])
@pytest.mark.parametrize("shape,kind", [((3, 96, 72, 48, 48, 3, 1), 2), ((2, 48, 36, 64, 64, 3, 1), 2), ((3, 24, 18, 48, 64, 1, 1), 1),
                                        ((2, 48, 36, 48, 48, 3, 2), 1)], ids=str)
def test_opt_in_epilogues_match_cpu(shape, kind, epi):
    """The opt-in epilogue variants (HRNET_TUNE_EPILOGUE = coal: warp-staged coalesced stores; batch: all TMEM loads of a
    <= 64-channel tile before one wait) run on hardware here, so selecting them never executes unvalidated code."""
    case = G.conv_case(*shape, relu=True, residual=(shape[6] == 1 and shape[3] == shape[4]), out_f32=False, seed=12)
    _lib.set_debug_tune({_lib.TUNE_EPILOGUE: _lib.EPI_COAL if epi == "coal" else _lib.EPI_BATCH})
    try:
        out = G.run_conv(case, use_tc=kind)
    finally:
        _lib.set_debug_tune(None)
    err = (out - case["ref"]).abs().max().item()
    assert err <= G.conv_tolerance(case), f"max abs err {err}"


@pytest.mark.parametrize("shape", [(2, 24, 18, 192, 48, 1, 1), (2, 48, 36, 96, 192, 3, 2)], ids=str)
def test_conv_tc_f32_output_no_relu(shape):
    case = G.conv_case(*shape, relu=False, residual=False, out_f32=True, seed=2)
    out = G.run_conv(case, use_tc=True)
    assert (out - case["ref"]).abs().max().item() <= 1e-3 * max(1.0, case["ref"].abs().max().item())
    assert (out < 0).any()


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64, 3, 1), (2, 24, 18, 48, 96, 3, 2), (1, 12, 9, 64, 24, 1, 1)], ids=str)
def test_conv_simt_matches_cpu(shape):
    case = G.conv_case(*shape, relu=True, residual=False, out_f32=False, seed=3)
    out = G.run_conv(case, use_tc=False)
    assert (out - case["ref"]).abs().max().item() <= G.conv_tolerance(case)


def test_conv_tc_full_size_linearity():
    """BASELINE-size property test (W48 branch 2 at N=64, no oracle needed): conv is linear, so with
    scale=1, bias=0, no ReLU, fp32 output: conv(2x) == 2*conv(x) exactly (power-of-two scaling)."""
    d = torch.device("cuda")
    g = torch.Generator().manual_seed(5)
    n, h, w, c = 64, 24, 18, 192
    x = (torch.randn(n, h, w, c, generator=g) * 0.25).to(torch.float16).to(d)
    wt = (torch.randn(c, 3, 3, c, generator=g) / (9 * c) ** 0.5).to(torch.float16).to(d)
    one, zero = torch.ones(c, device=d), torch.zeros(c, device=d)
    outs = []
    for xin in (x, x * 2):
        o = torch.empty(n, h, w, c, dtype=torch.float32, device=d)
        _lib.check(G.lib().hrnet_conv_bn_act(G.ptr(xin), G.ptr(wt), G.ptr(one), G.ptr(zero), None, G.ptr(o), n, h, w,
                                             c, c, 3, 1, 0, 1, 1, G.stream()))
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[1], outs[0] * 2)
    assert outs[0].abs().max() > 0.1


def test_fuse_kernel():
    d = torch.device("cuda")
    g = torch.Generator().manual_seed(0)
    n, h, w, c = 2, 16, 8, 48
    s0 = torch.randn(n, h, w, c, generator=g).to(torch.float16)
    s1 = torch.randn(n, h // 2, w // 2, c, generator=g)
    s2 = torch.randn(n, h // 4, w // 4, c, generator=g).to(torch.float16)
    s3 = torch.randn(n, h // 8, w // 8, c, generator=g)
    srcs = [s0, s1, s2, s3]
    dev = [t.to(d) for t in srcs]
    up = lambda t, k: t.float().repeat_interleave(2 ** k, dim=1).repeat_interleave(2 ** k, dim=2)
    ref = up(s0, 0)
    for k in (1, 2, 3):
        ref = ref + up(srcs[k], k)
    ref = torch.relu(ref).to(torch.float16)
    out = torch.empty(n, h, w, c, dtype=torch.float16, device=d)
    arr = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in dev])
    shifts = (ctypes.c_int * 4)(0, 1, 2, 3)
    f32 = (ctypes.c_int * 4)(0, 1, 0, 1)
    _lib.check(G.lib().hrnet_fuse(arr, shifts, f32, 4, G.ptr(out), n, h, w, c, 1, G.stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)


def _argmax_gpu(hm, boxes):
    d = torch.device("cuda")
    n, J, Hh, Wh = hm.shape
    hmd = torch.from_numpy(hm).to(d)
    bx = torch.from_numpy(boxes).to(d) if boxes is not None else None
    joints = torch.empty(n, J, 3, dtype=torch.float32, device=d)
    idx = torch.empty(n, J, dtype=torch.int32, device=d)
    _lib.check(G.lib().hrnet_argmax(G.ptr(hmd), n, J, Hh, Wh, G.ptr(bx), G.ptr(joints), G.ptr(idx), G.stream()))
    torch.cuda.synchronize()
    return joints.cpu().numpy(), idx.cpu().numpy()


def test_argmax_edge_fixture_bit_exact(golden_dir):
    """Fed the reference's own fp32 heat-maps: indices, confidences and (y, x) are bit-identical,
    incl. exact ties (first occurrence), all-equal maps, negative maps and non-integer box scaling."""
    g = np.load(os.path.join(golden_dir, "decode_edge.npz"))
    joints, idx = _argmax_gpu(g["heatmaps"], g["boxes"])
    assert np.array_equal(idx, g["argmax"].astype(np.int32))
    assert np.array_equal(joints.view(np.uint32), g["pts"].view(np.uint32))


def test_argmax_reference_heatmaps_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "w48_384x288_n1_default.npz"))
    joints, idx = _argmax_gpu(g["heatmaps"], None)
    assert np.array_equal(idx, g["argmax"].astype(np.int32))
    assert np.array_equal(joints.view(np.uint32), g["pts"].view(np.uint32))


def test_argmax_random_ties_nan_and_full_size():
    rng = np.random.default_rng(3)
    hm = rng.integers(0, 4, (64, 17, 96, 72)).astype(np.float32)      # BASELINE size; heavy ties
    hm[5, 2, 40, 10] = np.nan
    hm[5, 2, 50, 10] = np.nan
    boxes = np.concatenate([rng.uniform(0, 50, (64, 2)), rng.uniform(60, 400, (64, 2))], axis=1).astype(np.float32)
    joints, idx = _argmax_gpu(hm, boxes)
    pts, oidx = O.decode_joints(hm, boxes)
    assert np.array_equal(idx, oidx.astype(np.int32))
    assert np.array_equal(joints.view(np.uint32), pts.view(np.uint32))
    assert idx[5, 2] == 40 * 72 + 10   # np.argmax: first NaN wins


def test_argmax_empty_batch_is_noop():
    _lib.check(G.lib().hrnet_argmax(None, 0, 17, 64, 48, None, None, None, G.stream()))


@pytest.mark.parametrize("shape", [(3, 300, 210, 256, 192), (2, 480, 640, 384, 288), (1, 30, 20, 384, 288), (2, 385, 289, 384, 288),
                                   (1, 700, 500, 257, 191)], ids=str)
def test_device_cubic_resize_equals_opencv_kernel(shape):
    """hrnet_resize_cubic_u8 (the resize of SimpleHRNet.py:216-220 on the device) == cv2.resize(INTER_CUBIC) evaluated by
    OpenCV's own kernel (vendor path off), bit for bit; and == its host restatement."""
    import cv2
    from simple_hrnet_b200.preprocess import CubicResizer, resize_cubic_reference
    n, sh, sw, dh, dw = shape
    img = np.random.default_rng(sh + sw).integers(0, 256, (n, sh, sw, 3), dtype=np.uint8)
    img[0, : sh // 3] = 255; img[0, sh // 3: 2 * sh // 3] = 0
    out = CubicResizer(torch.device("cuda:0"))(img, dh, dw).cpu().numpy()
    assert np.array_equal(out, resize_cubic_reference(img, dh, dw))
    was = cv2.useOptimized()
    cv2.setUseOptimized(False)
    try:
        want = np.stack([cv2.resize(i, (dw, dh), interpolation=cv2.INTER_CUBIC) for i in img])
    finally:
        cv2.setUseOptimized(was)
    assert np.array_equal(out, want)


def test_device_crops_equal_pillow_resize():
    """hrnet_crop_resize_bilinear_u8: `image[y1:y2, x1:x2]` (+ zero padding) -> `ToPILImage -> Resize((H, W))` of the
    reference's multi-person crop transform (SimpleHRNet.py:166-171, 244-276) on the device == Pillow, bit for bit; boxes
    of every aspect, up- and down-scaling, crops touching the frame border, several frames."""
    from PIL import Image
    from simple_hrnet_b200.preprocess import CropResizer
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (3, 480, 640, 3), dtype=np.uint8)
    frames[1, :160] = 255; frames[1, 160:320] = 0
    specs = [(0, 10, 20, 210, 420, 0, 0, 0, 0), (0, 0, 0, 640, 480, 0, 0, 0, 0), (1, 300, 100, 340, 400, 0, 0, 93, 92),
             (1, 100, 200, 500, 260, 236, 237, 0, 0), (2, 600, 440, 640, 480, 0, 0, 0, 0), (2, 5, 7, 28, 37, 0, 0, 0, 1),
             (0, 17, 33, 305, 417, 0, 0, 0, 0), (2, 0, 100, 639, 479, 3, 0, 0, 0)]
    dev = torch.from_numpy(frames).cuda()
    cr = CropResizer(torch.device("cuda:0"))
    for (oh, ow) in ((384, 288), (256, 192)):
        out = cr(dev, specs, oh, ow).cpu().numpy()
        for i, (f, x1, y1, x2, y2, pt, pb, pl, pr) in enumerate(specs):
            crop = np.pad(frames[f][y1:y2, x1:x2], ((pt, pb), (pl, pr), (0, 0)))
            want = np.asarray(Image.fromarray(np.ascontiguousarray(crop)).resize((ow, oh), Image.BILINEAR))
            assert np.array_equal(out[i], want), (i, oh, ow, int(np.abs(out[i].astype(int) - want.astype(int)).max()))
    with pytest.raises(ValueError):
        cr(dev, [(0, -3, 0, 10, 10, 0, 0, 0, 0)], 384, 288)
