import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine
sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
eng = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), flags=2 | 64)   # NO_GRAPH so the debug path runs
eng.load_state_dict(sd)
x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(1)).cuda()
eng.forward_decode(x); torch.cuda.synchronize()
os.environ["HRNET_B200_DBG_GROUP"] = "1"
eng.forward_decode(x); torch.cuda.synchronize()
