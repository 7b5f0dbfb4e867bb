"""Profiling target: ONE forward of the headline config (HRNet-W48 384x288, batch 64) between
cudaProfilerStart/Stop.  Used as:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_forward.py
A number printed under ncu is never a bench value."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O  # noqa: E402
from simple_hrnet_b200 import B200Engine, _lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tune = {int(k): int(v) for k, v in (kv.split("=") for kv in sys.argv[3].split(","))} if len(sys.argv) > 3 else None   # "20=2,18=1"
sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
eng = B200Engine("hrnet", 48, 17, (384, 288), n, torch.device("cuda:0"), flags=flags, tune=tune)
eng.load_state_dict(sd)
x = torch.randn(n, 3, 384, 288, generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(2):
    eng.forward_decode(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.forward_decode(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ops per forward:", eng.launch_count)
