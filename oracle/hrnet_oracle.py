"""
ORACLE (test infrastructure, NOT product code).

CPU fp32 restatement of the simple-HRNet hot path, used only as the parity
checker by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs.  Nothing under `simple-hrnet_b200/` may import this
package: the product path is CUDA-only and fails loudly without its extension.

What is restated (reference file:line, relative to stefanopini/simple-HRNet):

* `hrnet_forward`        <- models_/hrnet.py:157-189 (HRNet.forward),
                            models_/hrnet.py:55-71  (StageModule.forward),
                            models_/modules.py:20-40 (Bottleneck.forward),
                            models_/modules.py:56-72 (BasicBlock.forward)
* `poseresnet_forward`   <- models_/poseresnet.py:108-122
* `decode_joints`        <- SimpleHRNet.py:296-308 (dup 431-443)
* `preprocess`           <- SimpleHRNet.py:216-222, 350-366, 149-153
* `predict`              <- SimpleHRNet.py:174-210 (multiperson=False branches only)
* `get_max_preds`, `get_final_preds`, `inverse_affine`
                         <- misc/utils.py:125-182, :46-79, :115-122 (evaluation-side decode)
* `flip_average`         <- training/COCO.py:206-212, misc/utils.py:9-29 (flip test)
* `adapt_box`, `multiperson_crops`
                         <- SimpleHRNet.py:227-278, :166-171 (multi-person glue around the path)

The arithmetic itself lives in a third-party dependency that is not vendored in
the reference tree: PyTorch (requirements.txt `torch>=1.4.0`, unpinned; this
image has 2.11.0) -> aten conv2d / batch_norm(eval) / relu / upsample_nearest2d
/ conv_transpose2d / max_pool2d, and NumPy (`numpy>=1.16`, image 2.3.5) argmax /
unravel_index.  The restatement calls the same aten ops functionally
(`torch.nn.functional`) on CPU in fp32, so it is bit-identical to the
reference modules by construction; that is pinned by
`tests/golden/make_golden.py`, which runs the *real* reference classes
(`models_.hrnet.HRNet`, `SimpleHRNet.SimpleHRNet.predict`) in the build
container and commits their outputs, and by `tests/test_oracle.py`, which
replays the oracle against those fixtures.

The reference ships no tests / golden vectors for this path, so
the committed fixtures generated from the live reference are the pin.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # models_/hrnet.py:33,42,49,80 (explicit) and nn.BatchNorm2d default in modules.py

_BN_SUFFIXES = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")


# --------------------------------------------------------------------------------------
# Parameter inventory: names/shapes exactly as the reference's state_dict
# --------------------------------------------------------------------------------------

def _conv(keys, name, cout, cin, k, bias=False):
    keys.append((name + ".weight", (cout, cin, k, k), "conv"))
    if bias:
        keys.append((name + ".bias", (cout,), "conv_bias"))


def _bn(keys, name, ch):
    keys.append((name + ".weight", (ch,), "bn_gamma"))
    keys.append((name + ".bias", (ch,), "bn_beta"))
    keys.append((name + ".running_mean", (ch,), "bn_mean"))
    keys.append((name + ".running_var", (ch,), "bn_var"))
    keys.append((name + ".num_batches_tracked", (), "bn_count"))


def hrnet_param_spec(c: int, nof_joints: int) -> List[Tuple[str, tuple, str]]:
    """Ordered (key, shape, kind) list equal to `HRNet(c, nof_joints).state_dict()`.

    Follows the module construction order of models_/hrnet.py:75-155.
    """
    keys: List[Tuple[str, tuple, str]] = []
    _conv(keys, "conv1", 64, 3, 3); _bn(keys, "bn1", 64)
    _conv(keys, "conv2", 64, 64, 3); _bn(keys, "bn2", 64)
    # layer1: 4 x Bottleneck(->256); block 0 has the 1x1 downsample (hrnet.py:86-95)
    for b in range(4):
        cin = 64 if b == 0 else 256
        p = f"layer1.{b}"
        _conv(keys, p + ".conv1", 64, cin, 1); _bn(keys, p + ".bn1", 64)
        _conv(keys, p + ".conv2", 64, 64, 3); _bn(keys, p + ".bn2", 64)
        _conv(keys, p + ".conv3", 256, 64, 1); _bn(keys, p + ".bn3", 256)
        if b == 0:
            _conv(keys, p + ".downsample.0", 256, 64, 1); _bn(keys, p + ".downsample.1", 256)
    # transition1 (hrnet.py:98-109)
    _conv(keys, "transition1.0.0", c, 256, 3); _bn(keys, "transition1.0.1", c)
    _conv(keys, "transition1.1.0.0", 2 * c, 256, 3); _bn(keys, "transition1.1.0.1", 2 * c)

    def stage_module(prefix, stage, outb):
        for i in range(stage):
            w = c * (2 ** i)
            for k in range(4):
                p = f"{prefix}.branches.{i}.{k}"
                _conv(keys, p + ".conv1", w, w, 3); _bn(keys, p + ".bn1", w)
                _conv(keys, p + ".conv2", w, w, 3); _bn(keys, p + ".bn2", w)
        for i in range(outb):
            for j in range(stage):
                p = f"{prefix}.fuse_layers.{i}.{j}"
                if i == j:
                    continue
                if i < j:
                    _conv(keys, p + ".0", c * 2 ** i, c * 2 ** j, 1); _bn(keys, p + ".1", c * 2 ** i)
                else:
                    for k in range(i - j - 1):
                        _conv(keys, f"{p}.{k}.0", c * 2 ** j, c * 2 ** j, 3); _bn(keys, f"{p}.{k}.1", c * 2 ** j)
                    k = i - j - 1
                    _conv(keys, f"{p}.{k}.0", c * 2 ** i, c * 2 ** j, 3); _bn(keys, f"{p}.{k}.1", c * 2 ** i)

    stage_module("stage2.0", 2, 2)
    _conv(keys, "transition2.2.0.0", 4 * c, 2 * c, 3); _bn(keys, "transition2.2.0.1", 4 * c)
    for m in range(4):
        stage_module(f"stage3.{m}", 3, 3)
    _conv(keys, "transition3.3.0.0", 8 * c, 4 * c, 3); _bn(keys, "transition3.3.0.1", 8 * c)
    stage_module("stage4.0", 4, 4)
    stage_module("stage4.1", 4, 4)
    stage_module("stage4.2", 4, 1)
    _conv(keys, "final_layer", nof_joints, c, 1, bias=True)
    return keys


_RESNET_SPEC = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}


def poseresnet_param_spec(resnet_size: int, nof_joints: int) -> List[Tuple[str, tuple, str]]:
    """Ordered (key, shape, kind) list equal to `PoseResNet(resnet_size, nof_joints).state_dict()`
    (models_/poseresnet.py:17-106).  Only the Bottleneck sizes (50/101/152) work in the
    reference (modules.py:51 breaks 18/34), so only those are restated."""
    layers = _RESNET_SPEC[resnet_size]
    keys: List[Tuple[str, tuple, str]] = []
    _conv(keys, "conv1", 64, 3, 7); _bn(keys, "bn1", 64)
    inplanes = 64
    for li, (planes, nblocks) in enumerate(zip([64, 128, 256, 512], layers)):
        stride = 1 if li == 0 else 2
        for b in range(nblocks):
            p = f"layer{li + 1}.{b}"
            _conv(keys, p + ".conv1", planes, inplanes, 1); _bn(keys, p + ".bn1", planes)
            _conv(keys, p + ".conv2", planes, planes, 3); _bn(keys, p + ".bn2", planes)
            _conv(keys, p + ".conv3", planes * 4, planes, 1); _bn(keys, p + ".bn3", planes * 4)
            if b == 0 and (stride != 1 or inplanes != planes * 4):
                _conv(keys, p + ".downsample.0", planes * 4, inplanes, 1); _bn(keys, p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    for d in range(3):
        # ConvTranspose2d weight layout is [Cin, Cout, kh, kw] (poseresnet.py:91-99)
        keys.append((f"deconv_layers.{3 * d}.weight", (inplanes, 256, 4, 4), "deconv"))
        _bn(keys, f"deconv_layers.{3 * d + 1}", 256)
        inplanes = 256
    _conv(keys, "final_layer", nof_joints, 256, 1, bias=True)
    return keys


def make_state_dict(spec: Sequence[Tuple[str, tuple, str]], seed: int = 0, bn: str = "random",
                    gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic synthetic weights for a parameter spec (there is no network for
    the official checkpoints).  Each tensor draws from its own generator seeded by
    (seed, index), so the result does not depend on the reference's constructor RNG
    order and is reproducible on the GPU box where /root/reference does not exist.

    conv: U(-b, b), b = gain/sqrt(fan_in) (PyTorch's default conv init bound).
    bn="default": gamma=1, beta=0, mean=0, var=1 (nn.BatchNorm2d defaults -> BN ~ identity).
    bn="random" : gamma~U(.5,1.1), beta~N(0,.05), mean~N(0,.05), var~U(.8,1.6) so a wrong
                  scale/bias fold is visible while activations stay O(1)
                  through the ~60-deep residual chain (heatmaps within about +-1.2).
    """
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for idx, (name, shape, kind) in enumerate(spec):
        g = torch.Generator().manual_seed(seed * 1_000_003 + idx)
        if kind in ("conv", "deconv"):
            if kind == "conv":
                fan_in = shape[1] * shape[2] * shape[3]
            else:  # ConvTranspose2d: PyTorch computes fan_in from dim 1
                fan_in = shape[1] * shape[2] * shape[3]
            b = gain / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif kind == "conv_bias":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.1
        elif kind == "bn_gamma":
            t = torch.ones(shape) if bn == "default" else 0.5 + 0.6 * torch.rand(shape, generator=g)
        elif kind in ("bn_beta", "bn_mean"):
            t = torch.zeros(shape) if bn == "default" else 0.05 * torch.randn(shape, generator=g)
        elif kind == "bn_var":
            t = torch.ones(shape) if bn == "default" else 0.8 + 0.8 * torch.rand(shape, generator=g)
        elif kind == "bn_count":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise ValueError(kind)
        sd[name] = t.contiguous()
    return sd


# --------------------------------------------------------------------------------------
# Forward restatement
# --------------------------------------------------------------------------------------

def _cbn(sd, x, conv, bn, stride=1, padding=0, relu=False):
    x = F.conv2d(x, sd[conv + ".weight"], sd.get(conv + ".bias"), stride=stride, padding=padding)
    x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     training=False, eps=BN_EPS)
    return F.relu(x) if relu else x


def _bottleneck(sd, x, p, stride=1, has_down=False):
    """models_/modules.py:20-40.  Stride sits on the 3x3 conv; the downsample is 1x1 stride s."""
    out = _cbn(sd, x, p + ".conv1", p + ".bn1", relu=True)
    out = _cbn(sd, out, p + ".conv2", p + ".bn2", stride=stride, padding=1, relu=True)
    out = _cbn(sd, out, p + ".conv3", p + ".bn3")
    res = _cbn(sd, x, p + ".downsample.0", p + ".downsample.1", stride=stride) if has_down else x
    return F.relu(out + res)


def _basic_block(sd, x, p):
    """models_/modules.py:56-72."""
    out = _cbn(sd, x, p + ".conv1", p + ".bn1", padding=1, relu=True)
    out = _cbn(sd, out, p + ".conv2", p + ".bn2", padding=1)
    return F.relu(out + x)


def _stage_module(sd, xs, prefix, stage, outb):
    """models_/hrnet.py:55-71: branches, then x_fused[i] = sum_j fuse[i][j](x[j]) (j ascending), ReLU."""
    assert len(xs) == stage
    ys = []
    for i, x in enumerate(xs):
        for k in range(4):
            x = _basic_block(sd, x, f"{prefix}.branches.{i}.{k}")
        ys.append(x)
    fused = []
    for i in range(outb):
        acc = None
        for j in range(stage):
            p = f"{prefix}.fuse_layers.{i}.{j}"
            if i == j:
                t = ys[j]
            elif i < j:
                t = _cbn(sd, ys[j], p + ".0", p + ".1")
                t = F.interpolate(t, scale_factor=2.0 ** (j - i), mode="nearest")  # hrnet.py:34
            else:
                t = ys[j]
                for k in range(i - j - 1):
                    t = _cbn(sd, t, f"{p}.{k}.0", f"{p}.{k}.1", stride=2, padding=1, relu=True)
                k = i - j - 1
                t = _cbn(sd, t, f"{p}.{k}.0", f"{p}.{k}.1", stride=2, padding=1)
            acc = t if acc is None else acc + t
        fused.append(F.relu(acc))
    return fused


@torch.no_grad()
def hrnet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, return_taps: bool = False):
    """models_/hrnet.py:157-189.  x: [N,3,H,W] fp32 -> heatmaps [N,J,H/4,W/4] fp32.

    With return_taps=True also returns a dict of intermediate activations (NCHW fp32)
    used by the per-layer parity tests to localise drift."""
    taps = {}
    x = _cbn(sd, x, "conv1", "bn1", stride=2, padding=1, relu=True)
    taps["stem1"] = x
    x = _cbn(sd, x, "conv2", "bn2", stride=2, padding=1, relu=True)
    taps["stem2"] = x
    for b in range(4):
        x = _bottleneck(sd, x, f"layer1.{b}", has_down=(b == 0))
    taps["layer1"] = x
    xs = [_cbn(sd, x, "transition1.0.0", "transition1.0.1", padding=1, relu=True),
          _cbn(sd, x, "transition1.1.0.0", "transition1.1.0.1", stride=2, padding=1, relu=True)]
    taps["trans1"] = list(xs)
    xs = _stage_module(sd, xs, "stage2.0", 2, 2)
    taps["stage2"] = list(xs)
    xs = [xs[0], xs[1], _cbn(sd, xs[-1], "transition2.2.0.0", "transition2.2.0.1", stride=2, padding=1, relu=True)]
    for m in range(4):
        xs = _stage_module(sd, xs, f"stage3.{m}", 3, 3)
        taps[f"stage3.{m}"] = list(xs)
    xs = [xs[0], xs[1], xs[2],
          _cbn(sd, xs[-1], "transition3.3.0.0", "transition3.3.0.1", stride=2, padding=1, relu=True)]
    xs = _stage_module(sd, xs, "stage4.0", 4, 4)
    taps["stage4.0"] = list(xs)
    xs = _stage_module(sd, xs, "stage4.1", 4, 4)
    taps["stage4.1"] = list(xs)
    xs = _stage_module(sd, xs, "stage4.2", 4, 1)
    taps["stage4.2"] = list(xs)
    out = F.conv2d(xs[0], sd["final_layer.weight"], sd["final_layer.bias"])
    return (out, taps) if return_taps else out


@torch.no_grad()
def poseresnet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, resnet_size: int = 50) -> torch.Tensor:
    """models_/poseresnet.py:108-122."""
    layers = _RESNET_SPEC[resnet_size]
    x = _cbn(sd, x, "conv1", "bn1", stride=2, padding=3, relu=True)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    inplanes = 64
    for li, (planes, nblocks) in enumerate(zip([64, 128, 256, 512], layers)):
        stride = 1 if li == 0 else 2
        for b in range(nblocks):
            has_down = b == 0 and (stride != 1 or inplanes != planes * 4)
            x = _bottleneck(sd, x, f"layer{li + 1}.{b}", stride=stride if b == 0 else 1, has_down=has_down)
            inplanes = planes * 4
    for d in range(3):
        x = F.conv_transpose2d(x, sd[f"deconv_layers.{3 * d}.weight"], None, stride=2, padding=1, output_padding=0)
        bn = f"deconv_layers.{3 * d + 1}"
        x = F.relu(F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"],
                                sd[bn + ".bias"], training=False, eps=BN_EPS))
    return F.conv2d(x, sd["final_layer.weight"], sd["final_layer.bias"])


# --------------------------------------------------------------------------------------
# Decode + predict restatement
# --------------------------------------------------------------------------------------

def decode_joints(heatmaps: np.ndarray, boxes: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """SimpleHRNet.py:296-308.  heatmaps [n,J,Hh,Wh] f32, boxes [n,4] (x1,y1,x2,y2).

    Returns (pts [n,J,3] f32 as (y, x, confidence), flat argmax index [n,J] int64).
    np.argmax: flat row-major, first occurrence wins.  Coordinates are evaluated in
    float64 (np.int64 * python float -> f64; the f32 box difference promotes) and then
    stored to float32, exactly like the reference's scalar expression."""
    n, J, Hh, Wh = heatmaps.shape
    flat = heatmaps.reshape(n, J, Hh * Wh)
    idx = np.argmax(flat, axis=2)                                   # first occurrence
    r, c = np.unravel_index(idx, (Hh, Wh))
    pts = np.empty((n, J, 3), dtype=np.float32)
    dy = (boxes[:, 3] - boxes[:, 1])                                # box dtype arithmetic (f32 or i32)
    dx = (boxes[:, 2] - boxes[:, 0])
    y = r.astype(np.float64) * 1. / Hh * dy[:, None].astype(np.float64) + boxes[:, 1:2].astype(np.float64)
    x = c.astype(np.float64) * 1. / Wh * dx[:, None].astype(np.float64) + boxes[:, 0:1].astype(np.float64)
    pts[:, :, 0] = y
    pts[:, :, 1] = x
    pts[:, :, 2] = np.take_along_axis(flat, idx[:, :, None], axis=2)[:, :, 0]
    return pts, idx.astype(np.int64)


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess(images_bgr_u8: np.ndarray, resolution: Tuple[int, int], interpolation=None) -> torch.Tensor:
    """SimpleHRNet.py:216-222 / 350-366 with the transform of :149-153
    (cv2.resize -> BGR2RGB -> ToTensor (/255, HWC->CHW) -> Normalize).  [n,H,W,3] u8 -> [n,3,H,W] f32."""
    import cv2
    if interpolation is None:
        interpolation = cv2.INTER_CUBIC
    H, W = resolution
    out = torch.empty(images_bgr_u8.shape[0], 3, H, W)
    mean = torch.tensor(IMAGENET_MEAN).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(3, 1, 1)
    for i, img in enumerate(images_bgr_u8):
        img = cv2.resize(img, (W, H), interpolation=interpolation)
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
        t = torch.from_numpy(np.ascontiguousarray(img)).permute(2, 0, 1).to(torch.float32).div(255)
        out[i] = (t - mean) / std
    return out


def predict(sd, images_bgr_u8: np.ndarray, resolution: Tuple[int, int], model: str = "hrnet",
            resnet_size: int = 50):
    """SimpleHRNet.predict with multiperson=False, return_heatmaps=True, return_bounding_boxes=True.
    Accepts [H,W,3] or [n,H,W,3]; returns [heatmaps, boxes, pts] with the reference's shapes
    (single: pts [1,J,3]; batch: pts [n,1,J,3], SimpleHRNet.py:475)."""
    single = images_bgr_u8.ndim == 3
    imgs = images_bgr_u8[None] if single else images_bgr_u8
    if imgs.ndim != 4:
        raise ValueError('Wrong image format.')
    old_h, old_w = imgs.shape[1], imgs.shape[2]
    x = preprocess(imgs, resolution)
    fwd = hrnet_forward if model == "hrnet" else (lambda s, t: poseresnet_forward(s, t, resnet_size))
    hm = fwd(sd, x).numpy()
    boxes = np.repeat(np.asarray([[0, 0, old_w, old_h]], dtype=np.float32), len(imgs), axis=0)
    pts, _ = decode_joints(hm, boxes)
    if not single:
        pts = np.expand_dims(pts, axis=1)
    return [hm, boxes, pts]


# --------------------------------------------------------------------------------------
# Evaluation-side decode: get_max_preds / get_final_preds / flip test
# --------------------------------------------------------------------------------------

def get_max_preds(batch_heatmaps: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """misc/utils.py:125-151.  heatmaps [n,J,Hh,Wh] f32 -> (preds [n,J,2] f32 as (x, y) in heat-map pixels,
    maxvals [n,J,1] f32).  torch.max over the flattened map: first maximal index, a NaN counts as the maximum
    (== np.argmax); `idx % width` / `floor(idx / width)` are evaluated on float32 indices (exact below 2^24);
    joints whose maximum is not > 0 are zeroed (`preds *= pred_mask`)."""
    n, J, Hh, Wh = batch_heatmaps.shape
    flat = batch_heatmaps.reshape(n, J, Hh * Wh)
    idx = np.argmax(flat, axis=2)
    maxvals = np.take_along_axis(flat, idx[:, :, None], axis=2)
    idxf = idx.astype(np.float32)
    preds = np.zeros((n, J, 2), dtype=np.float32)
    preds[:, :, 0] = np.fmod(idxf, np.float32(Wh))
    preds[:, :, 1] = np.floor(idxf / np.float32(Wh))
    mask = (maxvals > 0.0).astype(np.float32)          # NaN > 0 is False, like torch.gt
    preds *= mask
    return preds, maxvals.astype(np.float32)


def inverse_affine(center, scale, pixel_std, output_size) -> np.ndarray:
    """misc/utils.py:46-79 with rot = 0, shift = 0, inv = 1 (the call made by transform_preds, :115-122): the 2x3
    float64 matrix that maps heat-map pixels back to image pixels.  cv2.getAffineTransform solves the 3-point system,
    so the same routine is called here on the same three point pairs."""
    import cv2
    scale = np.asarray(scale) if np.ndim(scale) else np.array([scale, scale])   # dtype kept: f32 * python float stays f32
    center = np.asarray(center)
    scale_tmp = scale * 1.0 * pixel_std
    src_w = scale_tmp[0]
    dst_w, dst_h = output_size[0], output_size[1]
    zero_shift = np.array([0, 0], dtype=np.float32)
    src_dir = [0 * 1.0 - (src_w * -0.5) * 0.0, 0 * 0.0 + (src_w * -0.5) * 1.0]    # get_dir([0, -src_w / 2], 0): sin 0 = 0, cos 0 = 1
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * zero_shift
    src[1, :] = center + src_dir + scale_tmp * zero_shift
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + np.array([0, dst_w * -0.5], np.float32)

    def third(a, b):
        d = a - b
        return b + np.array([-d[1], d[0]], dtype=np.float32)

    src[2:, :] = third(src[0, :], src[1, :])
    dst[2:, :] = third(dst[0, :], dst[1, :])
    return cv2.getAffineTransform(np.float32(dst), np.float32(src))


def get_final_preds(post_processing: bool, batch_heatmaps: np.ndarray, center, scale, pixel_std) -> Tuple[np.ndarray, np.ndarray]:
    """misc/utils.py:154-182: get_max_preds, optional quarter-pixel shift towards the higher neighbour (float32,
    `torch.sign`: 0 for 0 and for NaN), then the inverse affine of every joint evaluated in float64
    (`np.dot(t, [x, y, 1.])`) and stored to float32.  center/scale [n,2]; returns (preds [n,J,2] f32, maxvals [n,J,1])."""
    coords, maxvals = get_max_preds(batch_heatmaps)
    n, J, Hh, Wh = batch_heatmaps.shape
    if post_processing:
        for i in range(n):
            for p in range(J):
                hm = batch_heatmaps[i, p]
                px = int(math.floor(coords[i, p, 0] + 0.5))
                py = int(math.floor(coords[i, p, 1] + 0.5))
                if 1 < px < Wh - 1 and 1 < py < Hh - 1:
                    with np.errstate(invalid="ignore"):
                        diff = np.array([hm[py][px + 1] - hm[py][px - 1], hm[py + 1][px] - hm[py - 1][px]], dtype=np.float32)
                    sgn = np.where(diff > 0, np.float32(1), np.where(diff < 0, np.float32(-1), np.float32(0))).astype(np.float32)
                    coords[i, p] += sgn * np.float32(.25)
    preds = coords.copy()
    for i in range(n):
        t = inverse_affine(center[i], scale[i], pixel_std, [Wh, Hh])
        for p in range(J):
            new_pt = np.array([coords[i, p, 0], coords[i, p, 1], 1.]).T
            preds[i, p, 0:2] = np.dot(t, new_pt)[:2]
    return preds, maxvals


def flip_average(output: np.ndarray, output_flipped: np.ndarray, flip_pairs) -> np.ndarray:
    """training/COCO.py:206-212 + misc/utils.py:9-29: `(output + flip_back(output_flipped, pairs)) * 0.5` on
    [n,J,Hh,Wh] f32 maps: the flipped output is mirrored along x and its left/right joints swapped back."""
    fb = output_flipped[:, :, :, ::-1].copy()
    for a, b in flip_pairs:
        tmp = fb[:, a].copy()
        fb[:, a] = fb[:, b]
        fb[:, b] = tmp
    return (output + fb) * np.float32(0.5)


# --------------------------------------------------------------------------------------
# Multi-person glue: box adaptation + crop of SimpleHRNet._predict_single
# --------------------------------------------------------------------------------------

def adapt_box(x1: int, y1: int, x2: int, y2: int, resolution: Tuple[int, int]):
    """SimpleHRNet.py:244-276: pad (not enlarge) a detection so that it matches the network's aspect ratio.
    Integer box in, (x1_new, y1_new, x2_new, y2_new), np.pad tuple or None out."""
    correction_factor = resolution[0] / resolution[1] * (x2 - x1) / (y2 - y1)
    if correction_factor > 1:
        center = y1 + (y2 - y1) // 2
        length = int(round((y2 - y1) * correction_factor))
        y1_new = int(center - length // 2)
        y2_new = int(center + length // 2)
        pad = (int(abs(y1_new - y1))), int(abs(y2_new - y2))
        return (x1, y1_new, x2, y2_new), (pad, (0, 0), (0, 0))
    elif correction_factor < 1:
        center = x1 + (x2 - x1) // 2
        length = int(round((x2 - x1) * 1 / correction_factor))
        x1_new = int(center - length // 2)
        x2_new = int(center + length // 2)
        pad = (abs(x1_new - x1)), int(abs(x2_new - x2))
        return (x1_new, y1, x2_new, y2), ((0, 0), pad, (0, 0))
    return (x1, y1, x2, y2), None


def multiperson_crops(image_bgr_u8: np.ndarray, detections, resolution: Tuple[int, int]):
    """SimpleHRNet.py:227-278: detections (rows x1, y1, x2, y2, ...) -> (network inputs [m,3,H,W] f32, boxes [m,4] int32).
    Each crop is `image[y1:y2, x1:x2, ::-1]` (BGR -> RGB), zero-padded to the network aspect ratio and passed through
    ToPILImage -> Resize((H, W)) -> ToTensor -> Normalize (:166-171)."""
    from torchvision.transforms import transforms
    tf = transforms.Compose([transforms.ToPILImage(), transforms.Resize((resolution[0], resolution[1])),
                             transforms.ToTensor(),
                             transforms.Normalize(mean=list(IMAGENET_MEAN), std=list(IMAGENET_STD))])
    m = len(detections) if detections is not None else 0
    boxes = np.empty((m, 4), dtype=np.int32)
    images = torch.empty((m, 3, resolution[0], resolution[1]))
    for i in range(m):
        x1, y1, x2, y2 = [int(round(float(v))) for v in detections[i][:4]]
        box, pad = adapt_box(x1, y1, x2, y2, resolution)
        crop = image_bgr_u8[y1:y2, x1:x2, ::-1]
        if pad is not None:
            crop = np.pad(crop, pad)
        images[i] = tf(crop)
        boxes[i] = box
    return images, boxes
