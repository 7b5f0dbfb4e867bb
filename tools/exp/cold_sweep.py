"""Isolated branch convs with COLD operands (rotating through buffer sets larger than the 126 MB L2) vs L2-hot ones,
role timers on: separates "the data comes from HBM" from "the clocks / the rest of the forward" as the reason why the
convs take ~1.4x longer inside the network than alone.  Usage: python tools/exp/cold_sweep.py 2> gpurun_out/cold_sweep.log"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["HRNET_B200_DBG"] = "1"
from simple_hrnet_b200 import _lib
from tests import gpu_util as G

d = torch.device("cuda")
shapes = {"b0": (64, 96, 72, 48, 48, 3, 1, 1, 2), "b1": (64, 48, 36, 96, 96, 3, 1, 1, 2),
          "b2": (64, 24, 18, 192, 192, 3, 1, 1, 1), "b3": (64, 12, 9, 384, 384, 3, 1, 1, 1)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
for name, (n, h, w, ci, co, k, s, res, kern) in shapes.items():
    sets = []
    per_set = n * h * w * (ci + 2 * co) * 2 + co * k * k * ci * 2
    nsets = max(2, int(600e6 // per_set) + 1) if name in ("b0", "b1") else 4
    nsets = min(nsets, 12)
    for i in range(nsets):
        x = (torch.randn(n, h, w, ci, device=d) * 0.5).to(torch.float16)
        wt = (torch.randn(co, k, k, ci, device=d) / (k * k * ci) ** 0.5).to(torch.float16)
        r = torch.randn(n, h // s, w // s, co, device=d).to(torch.float16)
        out = torch.empty(n, h // s, w // s, co, dtype=torch.float16, device=d)
        sets.append((x, wt, r, out))
    sc, bi = torch.ones(co, device=d), torch.zeros(co, device=d)
    for mode in ("hot", "cold", "flushed"):
        print(f"case {name} {mode} ({nsets} sets of {per_set / 1e6:.0f} MB)", file=sys.stderr, flush=True)
        for it in range(6):
            x, wt, r, out = sets[0] if mode == "hot" else sets[it % nsets]
            if mode == "flushed":
                flush.zero_()          # 256 MB of writes: everything else leaves L2
                torch.cuda.synchronize()
            _lib.check(G.lib().hrnet_conv_bn_act(G.ptr(x), G.ptr(wt), G.ptr(sc), G.ptr(bi), G.ptr(r), G.ptr(out), n, h, w, ci, co, k, s, 1, 0, kern, G.stream()))
        torch.cuda.synchronize()
