"""What bounds the branch convs: SM-side work or a chip-wide resource (L2 -> SM bandwidth)?

Runs single branch convs (N = 64 crops, W48 shapes) with the per-CTA role timers on (HRNET_B200_DBG=1) while sweeping
  * HRNET_B200_GRID_CAP: fewer CTAs -> each does more tiles.  A kernel bound by SM-side work slows down in proportion,
    a kernel bound by a chip-wide resource does not;
  * HRNET_B200_CS=2 (CTA pairs): 35 % fewer L2 -> SM bytes for the same MMA work (each CTA loads half the weight tile);
  * HRNET_B200_M2 (two M-tiles per CTA sharing each weight stage): the same byte saving without clusters.
The kernel time is the `roles_done` max (globaltimer ns since the first CTA entry) of the [dbg-ns] line; the [dbg] line
gives cycles per role, so cycles / ns = the SM clock during the kernel.
Usage: python tools/exp/l2_sweep.py 2> gpurun_out/l2_sweep.log
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["HRNET_B200_DBG"] = "1"
from simple_hrnet_b200 import _lib
from tests import gpu_util as G

d = torch.device("cuda")
# n, h, w, cin, cout, k, stride, residual, kernel (1 = im2col, 2 = halo patch)
shapes = {"b2": (64, 24, 18, 192, 192, 3, 1, 1, 1), "b3": (64, 12, 9, 384, 384, 3, 1, 1, 1),
          "b0": (64, 96, 72, 48, 48, 3, 1, 1, 2), "b1": (64, 48, 36, 96, 96, 3, 1, 1, 2)}
sweeps = []
for name in ("b2", "b3"):
    for env in ({}, {"HRNET_B200_CS": "2"}):
        for cap in (148, 108, 74, 54):
            sweeps.append((name, dict(env, HRNET_B200_GRID_CAP=str(cap))))
for name in ("b0", "b1"):
    for cap in (148, 74):
        sweeps.append((name, {"HRNET_B200_GRID_CAP": str(cap)}))
if len(sys.argv) > 1:
    sweeps = [s for s in sweeps if s[0] in sys.argv[1:]]

data = {}
for name, (n, h, w, ci, co, k, s, res, kern) in shapes.items():
    x = (torch.randn(n, h, w, ci, device=d) * 0.5).to(torch.float16)
    wt = (torch.randn(co, k, k, ci, device=d) / (k * k * ci) ** 0.5).to(torch.float16)
    sc, bi = torch.ones(co, device=d), torch.zeros(co, device=d)
    r = torch.randn(n, h // s, w // s, co, device=d).to(torch.float16) if res else None
    out = torch.empty(n, h // s, w // s, co, dtype=torch.float16, device=d)
    data[name] = (x, wt, sc, bi, r, out)

keys = ("HRNET_B200_CS", "HRNET_B200_M2", "HRNET_B200_GRID_CAP")
for name, env in sweeps:
    for kk in keys:
        os.environ.pop(kk, None)
    os.environ.update(env)
    n, h, w, ci, co, k, s, res, kern = shapes[name]
    x, wt, sc, bi, r, out = data[name]
    print(f"case {name} {h}x{w} {ci}->{co} kernel={kern} env={env}", file=sys.stderr, flush=True)
    for _ in range(3):
        rc = G.lib().hrnet_conv_bn_act(G.ptr(x), G.ptr(wt), G.ptr(sc), G.ptr(bi), G.ptr(r), G.ptr(out), n, h, w, ci, co, k,
                                       s, 1, 0, kern, G.stream())
        if rc:
            print("  rc", rc, _lib.last_error() if hasattr(_lib, "last_error") else "", file=sys.stderr, flush=True)
            break
    torch.cuda.synchronize()
