"""Builds and runs tools/exp/mma_commit.cu: per-group cost of G MMAs + commit, with / without concurrent bulk copies."""
import ctypes, os, subprocess, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_commit.so")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "mma_commit.cu")):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                           "-o", so, os.path.join(HERE, "mma_commit.cu")])
if len(sys.argv) > 1 and sys.argv[1] == "build":
    sys.exit(0)
lib = ctypes.CDLL(so)
lib.exp_mma_commit.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(4, dtype=torch.int64, device="cuda")
src = torch.zeros(16 << 20, dtype=torch.uint8, device="cuda")
print("   N   G    R stages commit_every copy_kb | clk/group issue  clk/group done  clk/MMA  copies")
for (N, G, R, st, ce, ck) in [(64, 4, 256, 1, 0, 0), (64, 4, 256, 1, 1, 0), (64, 4, 256, 4, 1, 0), (64, 4, 256, 8, 1, 0),
                              (64, 4, 256, 4, 1, 16), (64, 4, 256, 4, 0, 16), (64, 4, 256, 4, 1, 8), (64, 4, 256, 4, 1, 4),
                              (48, 3, 256, 4, 1, 0), (48, 3, 256, 4, 1, 16), (48, 27, 64, 2, 1, 0), (48, 27, 64, 2, 1, 16),
                              (192, 4, 256, 4, 1, 0), (192, 4, 256, 4, 1, 16), (192, 4, 256, 4, 0, 16), (96, 4, 256, 4, 1, 0),
                              (96, 4, 256, 4, 1, 16), (256, 4, 128, 3, 1, 0), (256, 4, 128, 3, 1, 16)]:
    out.zero_()
    rc = lib.exp_mma_commit(N, G, R, st, ce, ck, src.data_ptr(), out.data_ptr())
    o = out.cpu().tolist()
    print(f"{N:4d} {G:3d} {R:4d} {st:6d} {ce:12d} {ck:7d} | {o[0] / R:15.1f} {o[1] / R:15.1f} {o[1] / R / G:8.1f} {o[2]:7d}", flush=True)
