"""Per-op output hashes of two forward(x[:k]) runs (HRNET_FLAG_NO_GRAPH + HRNET_B200_DBG_CHECKSUM): prints the ops whose
outputs differ between the runs.  Usage: python tools/dbg_checksum.py [k] [extra flags]"""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from oracle import hrnet_oracle as O
    from simple_hrnet_b200 import B200Engine
    k, flags = int(sys.argv[2]), int(sys.argv[3])
    sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
    e = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), flags=2 | flags)
    e.load_state_dict(sd)
    x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(3)).cuda()
    for rep in range(3):
        print(f"[run] {rep}", file=sys.stderr, flush=True)
        e.forward_decode(x[:k], return_heatmaps=True)
        torch.cuda.synchronize()
    sys.exit(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 5
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
env = dict(os.environ, HRNET_B200_DBG_CHECKSUM="1")
p = subprocess.run([sys.executable, __file__, "child", str(k), str(flags)], env=env, capture_output=True, text=True)
runs = []
for line in p.stderr.splitlines():
    if line.startswith("[run]"):
        runs.append([])
    elif line.startswith("[cs]") and runs:
        runs[-1].append(line.split()[1:])
print(f"k={k} flags={flags}: {len(runs)} runs, {len(runs[0]) if runs else 0} ops; rc={p.returncode}")
if p.returncode:
    print(p.stderr[-2000:])
for a in range(1, len(runs)):
    diff = [(i, n) for (i, n, h0), (_, _, h1) in zip(runs[0], runs[a]) if h0 != h1]
    print(f"run 0 vs run {a}: {len(diff)} ops differ; first: {diff[:12]}")
