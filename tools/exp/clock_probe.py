"""SM clock / power / throttle reasons sampled by nvidia-smi every 20 ms while the W48 forward runs back to back for ~3 s."""
import os, subprocess, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import B200Engine
sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(1)).cuda()
eng = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"))
eng.load_state_dict(sd)
for _ in range(5):
    eng.forward_decode(x)
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active",
                      "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
t0 = time.time()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
nit = 0
while time.time() - t0 < 3.0:
    for _ in range(20):
        eng.forward_decode(x)
    nit += 20
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
time.sleep(0.2)
p.terminate()
out = p.stdout.read().strip().splitlines()
print(f"{nit} forwards, {e0.elapsed_time(e1) / nit:.3f} ms each; {len(out)} samples")
for l in out[::4]:
    print(l)
