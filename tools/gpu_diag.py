"""First-contact diagnostics on a B200 (run under gpurun): per-shape conv parity for the tcgen05 and
SIMT kernels with failure dumps (gpurun_out/diag_*.npz) for offline analysis."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import gpu_util as G  # noqa: E402
from tests.test_gpu_kernels import CONV_SHAPES  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def main():
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    shapes = CONV_SHAPES if len(sys.argv) < 2 else CONV_SHAPES[: int(sys.argv[1])]
    for i, shape in enumerate(shapes):
        for use_tc in (False, True):
            tag = "tc" if use_tc else "simt"
            try:
                case = G.conv_case(*shape, relu=False, residual=False, out_f32=True, seed=7)
                t0 = time.time()
                out = G.run_conv(case, use_tc=use_tc)
                ref = case["ref"]
                diff = (out - ref).abs()
                bad = torch.isnan(out) | (diff > 2e-3 * max(1.0, ref.abs().max().item()))
                msg = f"[{i}] {shape} {tag}: max err {diff[~torch.isnan(diff)].max().item() if (~torch.isnan(diff)).any() else float('nan'):.3e} " \
                      f"nan {int(torch.isnan(out).sum())} bad {int(bad.sum())}/{bad.numel()} ({time.time() - t0:.2f}s)"
                print(msg, flush=True)
                if bad.any():
                    n, oh, ow, co = out.shape
                    badpix = bad.any(dim=3).reshape(-1)
                    badch = bad.reshape(-1, co).any(dim=0)
                    print(f"     bad pixels {int(badpix.sum())}/{badpix.numel()} first {badpix.nonzero()[:12].flatten().tolist()}"
                          f" bad channels {int(badch.sum())}/{co} first {badch.nonzero()[:12].flatten().tolist()}", flush=True)
                    np.savez_compressed(os.path.join(OUT, f"diag_{i}_{tag}.npz"), out=out.numpy(), ref=ref.numpy(),
                                        x=case["x"].numpy(), w=case["w"].numpy(), shape=np.asarray(shape))
            except Exception as e:  # noqa: BLE001
                print(f"[{i}] {shape} {tag}: EXCEPTION {e}", flush=True)
                traceback.print_exc()
                if "CUDA" in str(e) or "cuda" in str(e):
                    print("CUDA context likely poisoned; stopping diag", flush=True)
                    return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
