import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_mma.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "mma_rate.cu")])
lib = ctypes.CDLL(so)
lib.exp_mma_rate.argtypes = [ctypes.c_int] * 7 + [ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("N a_sbo_rows a_row_off d_col taps a_swizzle | issue cyc/MMA  complete cyc/MMA  (ideal math N/2)")
for (N, sbo, off, col, taps, sw) in ((96, 10, 0, 0, 0, 64), (96, 10, 11, 0, 0, 64), (96, 8, 0, 0, 0, 64), (48, 10, 3, 0, 0, 64), (192, 8, 0, 0, 0, 64)) + tuple(t + (128,) for t in ((48, 8, 0, 0, 0), (48, 10, 0, 0, 0), (48, 10, 11, 0, 0), (48, 10, 0, 48, 0), (48, 10, 0, 96, 0), (48, 10, 0, 64, 0),
                                 (48, 10, 0, 0, 1), (48, 10, 0, 48, 1), (48, 10, 0, 144, 1), (96, 10, 0, 0, 0), (96, 10, 0, 96, 0), (96, 10, 0, 288, 0),
                                 (192, 8, 0, 0, 0), (192, 8, 0, 192, 0), (32, 10, 0, 0, 0), (32, 10, 0, 32, 0), (64, 10, 0, 64, 0))):
    iters = 2700
    for _ in range(2):
        assert lib.exp_mma_rate(N, iters, sbo, off, col, taps, sw, out.data_ptr()) == 0
    a, b = out.tolist()
    print(f"{N:4d} {sbo:3d} {off:3d} {col:4d} {taps:2d} {sw:4d} | {a / iters:8.1f} {b / iters:8.1f}   ({N / 2})", flush=True)
