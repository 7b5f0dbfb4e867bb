"""TMA tiled-2D load rate vs global row pitch (contiguous rows vs 128 B slices of wider rows), see tma_bw.cu."""
import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_tma.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "tma_bw.cu")])
lib = ctypes.CDLL(so)
lib.exp_tma_bw.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p]
d = torch.device("cuda")
cyc = torch.zeros(148, dtype=torch.int64, device=d)
print("grid pitch_elems box_rows depth | cycles/load  cycles/row  B/clk/SM")
for pitch in (64, 192, 384, 1728, 3456):
    rows = (1 << 23) // pitch          # 16 MB matrix: L2 resident
    mat = torch.randn(rows, pitch, device=d).to(torch.float16)
    for grid in (1, 148):
        for box_rows in (128, 192):
            for depth in (4, 8):
                if depth * box_rows * 128 > 200 * 1024: continue
                iters = 400
                for _ in range(2):
                    assert lib.exp_tma_bw(mat.data_ptr(), rows, pitch, box_rows, 128, depth, iters, grid, cyc.data_ptr()) == 0
                c = cyc[:grid].max().item()
                print(f"{grid:4d} {pitch:5d} {box_rows:4d} {depth:2d} | {c / iters:9.1f} {c / iters / box_rows:7.2f} {box_rows * 128 * iters / c:8.2f}", flush=True)
