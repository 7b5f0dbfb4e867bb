import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_issue2.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "tma_issue2.cu")])
lib = ctypes.CDLL(so)
lib.exp_tma_issue_pitch.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 9 + [ctypes.c_void_p]
d = torch.device("cuda")
cyc = torch.zeros(148 * 8, dtype=torch.int64, device=d)
print("grid pitch nwarps box_rows k depth | cycles/step  cycles/row(SM)  B/clk/SM")
for pitch in (64, 192, 384, 1728):
    rows = (1 << 23) // pitch
    mat = torch.randn(rows, pitch, device=d).to(torch.float16)
    for grid in (1, 148):
        for nwarps, box_rows, k, depth in ((4, 64, 2, 2), (4, 32, 4, 2), (2, 128, 2, 2), (4, 128, 1, 3)):
            iters = 300
            rc = 0
            for _ in range(2):
                rc = lib.exp_tma_issue_pitch(mat.data_ptr(), rows, pitch, box_rows, k, depth, iters, grid, nwarps, 0, cyc.data_ptr())
            if rc == -5: continue
            assert rc == 0, rc
            c = cyc.view(148, 8)[:grid, :nwarps].max().item()
            nrows = nwarps * k * box_rows
            print(f"{grid:4d} {pitch:5d} {nwarps:2d} {box_rows:4d} {k:2d} {depth:2d} | {c / iters:9.1f} {c / iters / nrows:7.2f} {nrows * 128 * iters / c:8.2f}", flush=True)
