import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_mma.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "mma_rate.cu")])
lib = ctypes.CDLL(so)
lib.exp_mma_rate.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
print("N a_sbo_rows | issue cyc/MMA  complete cyc/MMA  (ideal math N/2)")
for sbo in (8, 10):
    for N in (16, 32, 48, 64, 96, 128, 192, 256):
        iters = 2000
        for _ in range(2):
            assert lib.exp_mma_rate(N, iters, sbo, out.data_ptr()) == 0
        a, b = out.tolist()
        print(f"{N:4d} {sbo:3d} | {a / iters:8.1f} {b / iters:8.1f}   ({N / 2})", flush=True)
