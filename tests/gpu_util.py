"""Helpers for the `-m gpu` parity tests: every CUDA call goes through the C ABI (ctypes)."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from simple_hrnet_b200 import _lib

_vp = ctypes.c_void_p


def lib():
    return _lib.load_library()


def ptr(t):
    return _vp(t.data_ptr()) if t is not None else None


def stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


def conv_case(n, ih, iw, cin, cout, k, stride, relu, residual, out_f32, seed=0):
    """Random conv problem.  Returns device NHWC fp16 inputs and the CPU fp32 reference computed on
    the same fp16-rounded operands (so only accumulation order differs)."""
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(n, ih, iw, cin, generator=g) * 0.5).to(torch.float16)
    w = (torch.randn(cout, k, k, cin, generator=g) / (k * k * cin) ** 0.5).to(torch.float16)
    scale = 0.5 + torch.rand(cout, generator=g)
    bias = 0.1 * torch.randn(cout, generator=g)
    oh, ow = ih // stride, iw // stride
    res = (torch.randn(n, oh, ow, cout, generator=g) * 0.5).to(torch.float16) if residual else None
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), stride=stride, padding=k // 2)
    y = y * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    if res is not None:
        y = y + res.float().permute(0, 3, 1, 2)
    if relu:
        y = F.relu(y)
    ref = y.permute(0, 2, 3, 1).contiguous()
    return dict(x=x, w=w, scale=scale, bias=bias, res=res, ref=ref, shape=(n, ih, iw, cin, cout, k, stride),
                relu=relu, out_f32=out_f32)


def run_conv(case, use_tc=True):
    n, ih, iw, cin, cout, k, stride = case["shape"]
    d = torch.device("cuda")
    x, w = case["x"].to(d), case["w"].to(d)
    scale, bias = case["scale"].to(d), case["bias"].to(d)
    res = case["res"].to(d) if case["res"] is not None else None
    out = torch.full((n, ih // stride, iw // stride, cout), float("nan"),
                     dtype=torch.float32 if case["out_f32"] else torch.float16, device=d)
    rc = lib().hrnet_conv_bn_act(ptr(x), ptr(w), ptr(scale), ptr(bias), ptr(res), ptr(out), n, ih, iw, cin, cout, k,
                                 stride, int(case["relu"]), int(case["out_f32"]), int(use_tc), stream())
    _lib.check(rc)
    torch.cuda.synchronize()
    return out.float().cpu()


def conv_tolerance(case):
    """fp16 output rounding (2^-11 relative) + fp32 accumulation-order noise."""
    return 2e-3 * max(1.0, case["ref"].abs().max().item())
