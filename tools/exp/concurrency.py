"""Do two persistent conv kernels with capped grids on two streams overlap?"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from simple_hrnet_b200 import _lib
from tests import gpu_util as G
d = torch.device("cuda")
def mk(c, h, w, n=64):
    x = (torch.randn(n, h, w, c, device=d) * 0.5).to(torch.float16)
    wt = (torch.randn(c, 3, 3, c, device=d) / (9 * c) ** 0.5).to(torch.float16)
    return dict(x=x, w=wt, sc=torch.ones(c, device=d), bi=torch.zeros(c, device=d),
                out=torch.empty(n, h, w, c, dtype=torch.float16, device=d), shape=(n, h, w, c))
def run(t, kern, stream):
    n, h, w, c = t["shape"]
    _lib.check(G.lib().hrnet_conv_bn_act(G.ptr(t["x"]), G.ptr(t["w"]), G.ptr(t["sc"]), G.ptr(t["bi"]), None, G.ptr(t["out"]),
                                         n, h, w, c, c, 3, 1, 1, 0, kern, ctypes.c_void_p(stream.cuda_stream)))
a, b = mk(48, 96, 72), mk(384, 12, 9)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
cap = os.environ.get("HRNET_B200_GRID_CAP")
print("grid cap", cap)
def only_a():
    with torch.cuda.stream(s1):
        for _ in range(8): run(a, 2, s1)
def only_b():
    with torch.cuda.stream(s2):
        for _ in range(8): run(b, 1, s2)
def both():
    with torch.cuda.stream(s1):
        for _ in range(8): run(a, 2, s1)
    with torch.cuda.stream(s2):
        for _ in range(8): run(b, 1, s2)
print(f"8x b0 on s1 alone: {timeit(only_a):.1f} us; 8x b3 on s2 alone: {timeit(only_b):.1f} us; both concurrently: {timeit(both):.1f} us")
