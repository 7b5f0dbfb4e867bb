"""Builds and runs tools/exp/mma_issue_overhead.cu (see the variants in its header)."""
import ctypes, os, subprocess, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_issue_ovh.so")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "mma_issue_overhead.cu")):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           "-o", so, os.path.join(HERE, "mma_issue_overhead.cu")])
if len(sys.argv) > 1 and sys.argv[1] == "build":
    sys.exit(0)
lib = ctypes.CDLL(so)
lib.exp_mma_issue_overhead.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p]
out = torch.zeros(4, dtype=torch.int64, device="cuda")
print("variant   N   G    R | clk/group issue  clk/group done  clk/MMA")
for (N, G, R) in [(64, 4, 256), (192, 4, 256), (48, 27, 64), (64, 1, 256)]:
    for v in range(8):
        for _ in range(2):
            out.zero_()
            rc = lib.exp_mma_issue_overhead(v, N, G, R, out.data_ptr())
        o = out.cpu().tolist()
        print(f"{v:7d} {N:4d} {G:3d} {R:4d} | {o[0] / R:15.1f} {o[1] / R:15.1f} {o[1] / R / G:8.1f}", flush=True)
