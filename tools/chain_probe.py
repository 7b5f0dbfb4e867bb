"""Branch-chain diagnostics on the headline config (W48 384x288, 64 crops):
  1. per-chain role timers (HRNET_TUNE_CHAIN_DEBUG) alone on the GPU and on the in-forward grid  -> stderr
  2. per-class serial times from hrnet_profile_ops
  3. forward time for several SM splits between the chains of a module (HRNET_TUNE_CHAIN_SHARE*) and without chains
Usage: python tools/chain_probe.py [debug] [shares "a,b,c,d" ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hrnet_oracle as O
from simple_hrnet_b200 import _lib
for _a in sys.argv[1:]:
    if _a.startswith("lib="):              # A/B builds of the library (development only): lib=<path of another .so>
        _p = os.path.join(ROOT, _a[4:])
        _lib.library_path = lambda _p=_p: _p
        from simple_hrnet_b200 import build as _b
        _b.needs_build = lambda: False
        print("library", _p)
from simple_hrnet_b200 import B200Engine

sd = O.make_state_dict(O.hrnet_param_spec(48, 17), seed=0, bn="default")
x = torch.randn(64, 3, 384, 288, generator=torch.Generator().manual_seed(1)).cuda()


def fwd_ms(eng, reps=8, groups=6):
    """median over `groups` timings of `reps` back-to-back forwards (a single short timing moves +-3 % with the clocks)"""
    for _ in range(4):
        eng.forward_decode(x)
    torch.cuda.synchronize()
    ts = []
    for _ in range(groups):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.forward_decode(x)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    ts.sort()
    return 0.5 * (ts[len(ts) // 2] + ts[(len(ts) - 1) // 2])


def fwd_ms_dual(eng_a, eng_b, reps=8, groups=6):
    """two forwards in flight: steps alternate between two engines (same weights, own workspaces) on two streams"""
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    main = torch.cuda.current_stream()
    def go(n):
        for i in range(n):
            with torch.cuda.stream(sa if i % 2 == 0 else sb):
                (eng_a if i % 2 == 0 else eng_b).forward_decode(x)
    go(4); torch.cuda.synchronize()
    ts = []
    for _ in range(groups):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main); sa.wait_event(e0); sb.wait_event(e0)
        go(2 * reps)
        main.wait_stream(sa); main.wait_stream(sb); e1.record(main); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / (2 * reps))
    ts.sort()
    return 0.5 * (ts[len(ts) // 2] + ts[(len(ts) - 1) // 2])


args = sys.argv[1:]
base_tune = {_lib.TUNE_CHAIN_M2: 1} if "m1" in args else {}
if "pair" in args:
    base_tune[_lib.TUNE_CHAIN_PAIR] = 2
    print("im2col chains: CTA pairs (cta_group::2)")
for a in args:
    if a.startswith("t") and "=" in a:          # any knob: t<index>=<value>
        base_tune[int(a[1:a.index("=")])] = int(a[a.index("=") + 1:])
        print("tune", a)
    if a.startswith("stages"):
        base_tune[_lib.TUNE_CHAIN_STAGES] = int(a[6:])
        print("im2col chains:", a)
    if a.startswith("skip"):
        base_tune[_lib.TUNE_CHAIN_SKIP] = int(a[4:])
        print("im2col chains: epilogue experiment", a, "(results invalid)")
if "m1" in args:
    print("im2col chains: one M-tile per ticket")
if "debug" in args:
    eng = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), tune={**base_tune, _lib.TUNE_CHAIN_DEBUG: 1})
    eng.load_state_dict(sd)
    ops, desc = eng.profile_ops(x, iters=3)
    cls = {}
    for name, op, us in ops:
        if op["kind"] == 1 and ".branches." in name:
            key = (name.split(".")[0], op["cin"])
            cls.setdefault(key, []).append(us)
    for k, v in sorted(cls.items()):
        print(f"{k[0]} C={k[1]:3d}: {len(v):3d} convs, {sum(v) / len(v):6.1f} us per conv (chain alone on the GPU)")
    print(f"forward (default split): {fwd_ms(eng):.3f} ms")
    eng.close()
if "dual" in args:
    ea = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), tune=dict(base_tune))
    ea.load_state_dict(sd)
    eb = ea.clone_shared()
    print(f"one forward at a time: {fwd_ms(ea):.3f} ms/forward; two in flight on two streams: {fwd_ms_dual(ea, eb):.3f} ms/forward", flush=True)
    eb.close(); ea.close()
splits = [a for a in args if "," in a]
for sp in splits:
    t = {**base_tune, **{i: int(v) for i, v in enumerate(sp.split(","))}}
    eng = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), tune=t)
    eng.load_state_dict(sd)
    d = eng.describe()
    grids = [(c["module"], c["branch"], c["grid"]) for c in d["chains"] if c["module"] in (1, 2, 6)]
    print(f"split {sp}: {fwd_ms(eng):.3f} ms/forward   grids (module, branch, ctas): {grids}", flush=True)
    eng.close()
if "nochain" in args:
    for name, fl, tn in (("chains + exchange units per source branch (default)", 0, {}),
                         ("chains, every fuse-layer conv its own launch", 0, {_lib.TUNE_XUNIT: 3}),
                         ("chains + exchange unit per module with sum tickets", 0, {_lib.TUNE_XUNIT: 2}),
                         ("every conv its own launch", _lib.FLAG_NO_CHAIN, {_lib.TUNE_XUNIT: 3})):
        eng = B200Engine("hrnet", 48, 17, (384, 288), 64, torch.device("cuda:0"), flags=fl, tune={**base_tune, **tn})
        eng.load_state_dict(sd)
        print(f"{name}: {fwd_ms(eng):.3f} ms/forward ({eng.launch_count} launches)", flush=True)
        eng.close()
