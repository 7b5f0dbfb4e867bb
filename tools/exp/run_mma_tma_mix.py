"""MMA stream + concurrent TMA refills of its operand stages (see mma_tma_mix.cu): clk per 64-channel k-block."""
import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_mix.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "mma_tma_mix.cu")])
lib = ctypes.CDLL(so)
lib.exp_mix.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 7 + [ctypes.c_void_p]
d = torch.device("cuda")
rows = 1 << 18                       # 262144 rows x 64 fp16 = 32 MB: L2 resident
mat = (torch.randn(rows, 64, device=d) * 0.01).to(torch.float16)
act = (torch.randn(64, 18, 24, 192, device=d) * 0.01).to(torch.float16)
wgt = (torch.randn(192, 1728, device=d) * 0.01).to(torch.float16)
cyc = torch.zeros(148, dtype=torch.int64, device=d)
A = {0: "2d-random", 1: "im2col", 2: "4d-rows"}
print("grid  N_mma b_rows stages A-loads   B-loads pollers mode | clk/k-block (max CTA)  TMA B/clk/SM  clk/MMA")
for grid in (148, 31):
    for (n, b, stages) in ((192, 192, 4), (192, 192, 5), (192, 96, 4), (192, 96, 7), (96, 96, 7), (256, 256, 4), (256, 128, 4), (128, 128, 4)):
        for amode, bmode, poll in ((0, 0, 0), (1, 1, 0), (2, 1, 0), (1, 1, 8), (1, 1, 2), (0, 0, 8)):
            if (amode or bmode or poll) and n != 192:
                continue
            for mode, name in ((1, "mma"), (2, "tma"), (3, "both")):
                if mode == 1 and (amode or bmode):
                    continue
                iters = 2160
                for _ in range(2):
                    rc = lib.exp_mix(mat.data_ptr(), act.data_ptr(), wgt.data_ptr(), rows, n, b, stages, iters,
                                     mode | (amode << 2) | (bmode << 4) | (poll << 8), grid, cyc.data_ptr())
                    if rc == -5:
                        break
                    assert rc == 0, rc
                if rc == -5:
                    continue
                c = cyc[:grid].max().item() / iters
                ab = 120 * 128 if amode == 2 else 16384
                print(f"{grid:4d} {n:5d} {b:5d} {stages:4d}   {A[amode]:9s} {'weights' if bmode else 'random ':7s} {poll:3d}     {name:4s} | {c:9.1f}   {(ab + b * 128) / c if mode & 2 else 0:8.1f}   {c / 4:7.1f}", flush=True)
