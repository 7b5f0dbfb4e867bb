import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_issue.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "tma_issue.cu")])
lib = ctypes.CDLL(so)
lib.exp_tma_issue.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_void_p]
d = torch.device("cuda")
rows = 1 << 17
mat = torch.randn(rows, 64, device=d).to(torch.float16)
cyc = torch.zeros(148 * 8, dtype=torch.int64, device=d)
print("grid nwarps box_rows k depth pf | cycles/step  cycles/TMA  B/clk/SM")
for grid in (1, 148):
    for nwarps in (1, 2, 4):
        for (box_rows, k) in ((16, 1), (16, 2), (16, 4), (16, 8), (64, 1), (64, 4), (128, 2)):
            for depth in (2, 4):
                for pf in (0,):
                    iters = 300
                    rc = 0
                    for _ in range(2):
                        rc = lib.exp_tma_issue(mat.data_ptr(), rows, box_rows, k, depth, iters, grid, nwarps, pf, cyc.data_ptr())
                    if rc == -5:
                        continue
                    assert rc == 0, rc
                    c = cyc.view(148, 8)[:grid, :nwarps].max().item()
                    print(f"{grid:4d} {nwarps:2d} {box_rows:4d} {k:2d} {depth:2d} {pf} | {c / iters:9.1f} {c / iters / k:9.1f} "
                          f"{nwarps * k * box_rows * 128 * iters / c:8.2f}", flush=True)
