"""Profiling target: the stage-4 3x3 branch convs (N=64) launched one by one through the C ABI.
  ncu --set full --clock-control none --import-source on -k regex:conv_igemm -o gpurun_out/prof_convs \
      python tools/profile_convs.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_hrnet_b200 import _lib  # noqa: E402
from tests import gpu_util as G  # noqa: E402

shapes = [(48, 96, 72), (96, 48, 36), (192, 24, 18), (384, 12, 9)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
d = torch.device("cuda")
for (c, h, w) in shapes:
    x = (torch.randn(n, h, w, c, device=d) * 0.5).to(torch.float16)
    wt = (torch.randn(c, 3, 3, c, device=d) / (9 * c) ** 0.5).to(torch.float16)
    sc, bi = torch.ones(c, device=d), torch.zeros(c, device=d)
    res = torch.randn(n, h, w, c, device=d).to(torch.float16)
    out = torch.empty(n, h, w, c, dtype=torch.float16, device=d)
    for _ in range(reps):
        _lib.check(G.lib().hrnet_conv_bn_act(G.ptr(x), G.ptr(wt), G.ptr(sc), G.ptr(bi), G.ptr(res), G.ptr(out), n, h, w,
                                             c, c, 3, 1, 1, 0, 2 if c <= 96 else 1, G.stream()))
    torch.cuda.synchronize()
print("done")
