"""Times one forward (W48 384x288, batch 64) under different plan flags / env toggles."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine, _lib

sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(1)).cuda()
names = {0: "default", 32: "PARTITION", 2: "NO_GRAPH", 34: "NO_GRAPH|PARTITION", 8: "SERIAL", 16: "NO_PATCH", 4: "FUSE_F32", 64: "GROUP", 66: "GROUP|NO_GRAPH"}
flags_list = [int(a) for a in sys.argv[1:]] or [0, 32, 2, 34, 8]
for flags in flags_list:
    eng = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), flags=flags)
    eng.load_state_dict(sd)
    for _ in range(3):
        eng.forward_decode(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward_decode(x)
    e1.record(); torch.cuda.synchronize()
    print(f"flags={flags:3d} {names.get(flags, ''):24s} PDL={'off' if os.environ.get('HRNET_B200_NO_PDL') else 'on'}: {e0.elapsed_time(e1) / 10:.3f} ms/forward", flush=True)
    eng.close()
