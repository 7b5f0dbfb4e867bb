"""In-situ wall time of a stage-3 and a stage-4 StageModule's branch chains (launched together on their in-forward grids,
hrnet_profile_ops) for several SM splits (HRNET_TUNE_CHAIN_SHARE0..3, per mille).  The forward time moves +-2 % with the
clocks; the module times are what the split changes.   python tools/split_sweep.py [pair] a,b,c,d ..."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine, _lib

sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(1)).cuda()
base = {_lib.TUNE_CHAIN_PAIR: 2} if "pair" in sys.argv else {}
for sp in [a for a in sys.argv[1:] if "," in a] or ["0,0,0,0"]:
    t = dict(base)
    if sp != "0,0,0,0":
        t.update({i: int(v) for i, v in enumerate(sp.split(","))})
    eng = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), tune=t)
    eng.load_state_dict(sd)
    best = {}
    for rep in range(3):
        ops, _ = eng.profile_ops(x, iters=4)
        mods = {}
        for name, op, us in ops:
            if op["kind"] == 1 and ".branches." in name:
                mods[name.split(".branches.")[0]] = mods.get(name.split(".branches.")[0], 0.0) + us
        for k, v in mods.items():
            best[k] = min(best.get(k, 1e9), v)
    d = eng.describe()
    grids = {m: [c["grid"] for c in d["chains"] if c["module"] == m] for m in (1, 2, 6)}
    s2, s3, s4 = best["stage2.0"], sum(best[f"stage3.{i}"] for i in range(4)) / 4, sum(best[f"stage4.{i}"] for i in range(3)) / 3
    print(f"split {sp:>16s}: stage2 {s2:6.1f} us  stage3 {s3:6.1f} us  stage4 {s4:6.1f} us per module | all chains {s2 + 4 * s3 + 3 * s4:7.1f} us | grids {grids}", flush=True)
    eng.close()
