"""Batch invariance at the other BASELINE configs: HRNet-W32 256x192 (max batch 32) and PoseResNet-50 256x192 (max batch 64):
forward(x[:k]) vs forward(x)[:k] on the heat-maps, two runs per k."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine
t0 = time.time()
for arch, c, B, ks in (("hrnet", 32, 32, (3, 8, 16)), ("poseresnet", 50, 64, (3, 8, 33))):
    spec = O.hrnet_param_spec(c, 17) if arch == "hrnet" else O.poseresnet_param_spec(c, 17)
    sd = O.make_state_dict(spec, seed=0, bn="default")
    e = B200Engine(arch, c, 17, (256, 192), B, torch.device("cuda:0"))
    e.load_state_dict(sd)
    x = torch.randn(B, 3, 256, 192, generator=torch.Generator().manual_seed(3)).cuda()
    _, _, hB = e.forward_decode(x, return_heatmaps=True)
    hB = hB.clone()
    _, _, hB2 = e.forward_decode(x, return_heatmaps=True)
    out = [f"full batch repeatable {torch.equal(hB, hB2)}"]
    for k in ks:
        worst = 0.0
        for _ in range(2):
            _, _, hk = e.forward_decode(x[:k], return_heatmaps=True)
            worst = max(worst, float((hk - hB[:k]).abs().max()))
        out.append(f"k={k}: {worst:.2e}")
    print(f"{arch}-{c} B={B} [{time.time() - t0:.0f} s]: " + "; ".join(out), flush=True)
    e.close()
