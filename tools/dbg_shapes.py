"""Role timers (HRNET_TUNE_DEBUG via _lib.set_debug_tune) for arbitrary conv shapes: n,h,w,cin,cout,k,stride,residual,kernel"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_hrnet_b200 import _lib
from tests import gpu_util as G
d = torch.device("cuda")
shapes = [(64, 96, 72, 64, 256, 1, 1, 1, 1), (64, 96, 72, 256, 64, 1, 1, 0, 1), (64, 96, 72, 48, 48, 3, 2, 0, 1),
          (64, 96, 72, 48, 96, 3, 2, 0, 1), (64, 48, 36, 96, 48, 1, 1, 0, 1), (64, 192, 144, 64, 64, 3, 2, 0, 1),
          (64, 96, 72, 256, 48, 3, 1, 0, 2)]
if len(sys.argv) > 1:   # shapes as n,h,w,cin,cout,k,stride,residual,kernel ...
    shapes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for (n, h, w, ci, co, k, s, res, kern) in shapes:
    x = (torch.randn(n, h, w, ci, device=d) * 0.5).to(torch.float16)
    wt = (torch.randn(co, k, k, ci, device=d) / (k * k * ci) ** 0.5).to(torch.float16)
    sc, bi = torch.ones(co, device=d), torch.zeros(co, device=d)
    r = torch.randn(n, h // s, w // s, co, device=d).to(torch.float16) if res else None
    out = torch.empty(n, h // s, w // s, co, dtype=torch.float16, device=d)
    print(f"shape n={n} {h}x{w} {ci}->{co} k{k}s{s} res={res} kernel={kern}", file=sys.stderr, flush=True)
    for _ in range(2):
        _lib.check(G.lib().hrnet_conv_bn_act(G.ptr(x), G.ptr(wt), G.ptr(sc), G.ptr(bi), G.ptr(r), G.ptr(out), n, h, w, ci, co, k, s, 1, 0, kern, G.stream()))
    torch.cuda.synchronize()
