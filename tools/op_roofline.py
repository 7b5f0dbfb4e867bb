"""Per-op table for one serial forward (W48 384x288, batch 64): time, algorithmic GFLOP and HBM bytes, the two
roofline times (tensor: MEASURED_PEAKS sustained bf16, hbm: measured copy bandwidth) and the ratio to the larger one.
Groups identical shapes.  Usage: python tools/op_roofline.py [n] > gpurun_out/op_roofline.txt"""
import json, os, sys
from collections import OrderedDict
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
peaks = {"tf": 1450.0, "gbs": 6500.0}
try:
    mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peaks["tf"] = float(mp.get("bf16_tflops_sustained", peaks["tf"]))
    peaks["gbs"] = float(mp.get("hbm_gbs", peaks["gbs"]))
except Exception:
    pass
sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
eng = B200Engine("hrnet", 48, 17, (384, 288), n, torch.device("cuda:0"))
eng.load_state_dict(sd)
x = torch.randn(n, 3, 384, 288, generator=torch.Generator().manual_seed(1)).cuda()
ops, desc = eng.profile_ops(x, iters=5)
T = desc["tensors"]
groups = OrderedDict()
for name, op, us in ops:
    if op["kind"] == 0 or op["out"] < 0:   # conv kinds carry cin/cout; others: count bytes only
        pass
    to = T[op["out"]] if op["out"] >= 0 else None
    ti = T[op["in"]] if op["in"] >= 0 else None
    flops = 0.0; byts = 0.0
    esz = lambda t: 4 if t.get("f32", 0) == 1 else 2
    if to is not None:
        byts += n * to["H"] * to["W"] * to["C"] * esz(to)
    if ti is not None:
        byts += n * ti["H"] * ti["W"] * ti["C"] * esz(ti)
    if op["res"] >= 0:
        tr = T[op["res"]]; byts += n * tr["H"] * tr["W"] * tr["C"] * esz(tr)
    for s in op["src"][:op["nsrc"]]:
        if s >= 0:
            ts = T[s]; byts += n * ts["H"] * ts["W"] * ts["C"] * esz(ts)
    if op["cin"] and op["cout"] and to is not None:
        k = op["k"]
        flops = 2.0 * n * to["H"] * to["W"] * op["cout"] * op["cin"] * k * k
        byts += op["cout"] * op["cin"] * k * k * 2
    key = (op["kind"], op["cin"], op["cout"], op["k"], op["stride"], to["H"] if to else 0, to["W"] if to else 0,
           op["res"] >= 0, op["use_patch"], op["nsrc"], op["tc"].get("epi", 0) if op["use_tc"] else 0)
    g = groups.setdefault(key, {"names": [], "us": 0.0, "flops": flops, "bytes": byts})
    g["names"].append(name); g["us"] += us
print(f"peaks: {peaks['tf']:.0f} TF/s, {peaks['gbs']:.0f} GB/s; n={n}")
print(f"{'first op':40s} {'cnt':>3s} {'cin':>4s} {'cout':>4s} k s {'HxW':>7s} res pat {'us/op':>7s} {'GF':>6s} {'MB':>6s} {'t_tc':>6s} {'t_hbm':>6s} {'x roof':>6s} {'total us':>8s}")
tot = 0.0; tot_roof = 0.0
for key, g in sorted(groups.items(), key=lambda kv: -kv[1]["us"]):
    kind, cin, cout, k, s, H, W, res, pat, nsrc, epi = key
    cnt = len(g["names"]); us = g["us"] / cnt
    t_tc = g["flops"] / (peaks["tf"] * 1e12) * 1e6
    t_hbm = g["bytes"] / (peaks["gbs"] * 1e9) * 1e6
    roof = max(t_tc, t_hbm, 1e-9)
    tot += g["us"]; tot_roof += roof * cnt
    print(f"{g['names'][0][:40]:40s} {cnt:3d} {cin:4d} {cout:4d} {k} {s} {H:3d}x{W:<3d} {int(res):3d} {int(pat):3d} {us:7.1f} {g['flops']/1e9:6.1f} {g['bytes']/1e6:6.1f} {t_tc:6.1f} {t_hbm:6.1f} {us/roof:6.2f} {g['us']:8.1f}")
print(f"serial total {tot:.0f} us; sum of per-op rooflines {tot_roof:.0f} us")
