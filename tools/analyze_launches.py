"""Joins an ncu launch list of tools/profile_forward.py (run with flags NO_GRAPH|SERIAL = 10 so launches are in op
order) with the plan's op list and aggregates device time per layer class.
  python tools/analyze_launches.py gpurun_out/launches_serial.csv [n]"""
import collections
import csv
import re
import sys

sys.path.insert(0, ".")
from simple_hrnet_b200.engine import Plan  # noqa: E402

path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lines = [l for l in open(path) if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
times = [float(r["Metric Value"].replace(",", "")) / 1000 for r in rows]
names = [r["Kernel Name"].split("(")[0] for r in rows]
p = Plan("hrnet", 48, 17, (384, 288), n, flags=10)
d = p.describe()
ops = d["ops"]
assert len(ops) == len(times), (len(ops), len(times))
agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
for op, t, kn in zip(ops, times, names):
    nm = op["name"]
    tin = d["tensors"][op["in"]] if op["in"] >= 0 else None
    if op["kind"] == 1:
        flop = 2.0 * n * (tin["H"] // op["stride"]) * (tin["W"] // op["stride"]) * op["k"] ** 2 * op["cin"] * op["cout"]
        if ".branches." in nm:
            cls = f"branch conv C={op['cin']} {tin['H']}x{tin['W']}"
        elif "fuse_layers" in nm:
            cls = f"fuse conv k{op['k']}s{op['stride']} {op['cin']}->{op['cout']} @{tin['H']}x{tin['W']}"
        elif nm.startswith("layer1"):
            cls = f"layer1 k{op['k']} {op['cin']}->{op['cout']}"
        else:
            cls = f"{nm} k{op['k']}s{op['stride']} {op['cin']}->{op['cout']}"
        cls += " [patch]" if op["use_patch"] else f" [im2col cs{op['tc']['cs']}]"
    else:
        flop = 0.0
        cls = {0: "stem", 2: "fuse_sum", 3: "head", 4: "argmax"}[op["kind"]]
        if op["kind"] == 2:
            to = d["tensors"][op["out"]]
            cls += f" {to['C']}x{to['H']}x{to['W']} nsrc={op['nsrc']}"
    a = agg[cls]
    a[0] += t; a[1] += 1; a[2] += flop
T = sum(times)
print(f"total {T:.1f} us over {len(times)} launches (serialised, cold)")
for cls, (t, c, f) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    tf = f / t / 1e6 if t > 0 and f > 0 else 0
    print(f"{t:9.1f} us {100 * t / T:5.1f}%  x{c:3d}  avg {t / c:7.1f} us  {tf:7.1f} TFLOP/s  {cls}")
