"""Single-op reproduction of the staged-TMA-epilogue race (layer1.0.conv3 shape: 64 -> 256, 1x1, residual, N images):
run the conv several times, compare with the fp32 reference and between runs, report which rows / channel chunks differ."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from simple_hrnet_b200 import _lib
from tests import gpu_util as G
for fix in [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0").split(",")]:
    os.environ["HRNET_B200_EPIFIX"] = str(fix)
    for n in (2, 5, 16, 64):
        case = G.conv_case(n, 96, 72, 64, 256, 1, 1, relu=True, residual=True, out_f32=False, seed=1)
        ref = case["ref"].reshape(-1, 256)
        outs = [G.run_conv(case, use_tc=True).reshape(-1, 256) for _ in range(4)]
        tol = G.conv_tolerance(case)
        bad_rows = set()
        for o in outs:
            d = (o - ref).abs()
            bad_rows |= set(torch.nonzero(d.amax(dim=1) > tol).flatten().tolist())
        same = all(torch.equal(outs[0], o) for o in outs[1:])
        tiles = sorted({r // 128 for r in bad_rows})
        chunks = set()
        for o in outs:
            d = (o - ref).abs() > tol
            chunks |= set((torch.nonzero(d.any(dim=0)).flatten() // 64).tolist())
        print(f"fix={fix} n={n}: runs identical {same}; rows beyond tol {len(bad_rows)} in tiles {tiles[:20]}{'...' if len(tiles) > 20 else ''} "
              f"(of {ref.shape[0] // 128}); channel chunks {sorted(chunks)}", flush=True)
