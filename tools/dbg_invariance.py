"""Batch-invariance diagnostic: forward(x[:k]) vs forward(x)[:k] (heat-maps, W48 384x288) for several k, under the plan
flags given as argv[1] (int) and the kernel-variant env switches of the environment."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ks = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 5, 8, 16, 33, 64]
sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
e = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), flags=flags)
e.load_state_dict(sd)
x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(3)).cuda()
_, _, h64 = e.forward_decode(x, return_heatmaps=True)
h64 = h64.clone()
out = []
for k in ks:
    worst, same = 0.0, True
    prev = None
    for rep in range(3):
        _, _, hk = e.forward_decode(x[:k], return_heatmaps=True)
        worst = max(worst, float((hk - h64[:k]).abs().max()))
        if prev is not None and not torch.equal(prev, hk):
            same = False
        prev = hk.clone()
    out.append(f"k={k}: {worst:.2e}{'' if same else ' (reruns differ)'}")
print(f"flags={flags} env={ {k: v for k, v in os.environ.items() if k.startswith('HRNET_B200')} }: " + "; ".join(out), flush=True)
