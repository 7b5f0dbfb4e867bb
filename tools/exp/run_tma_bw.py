"""Per-SM TMA load throughput / latency sweep (see tma_bw.cu)."""
import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_tma.so")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                       "-o", so, os.path.join(HERE, "tma_bw.cu")])
lib = ctypes.CDLL(so)
lib.exp_tma_bw.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p]
d = torch.device("cuda")
rows = 1 << 17                       # 131072 rows x 64 fp16 = 16 MB: L2 resident
mat = torch.randn(rows, 64, device=d).to(torch.float16)
cyc = torch.zeros(148, dtype=torch.int64, device=d)
print("grid box_rows row_bytes depth | cycles/load  B/clk/SM  (max over CTAs)")
for grid in (1, 148):
    for (box_rows, row_bytes) in ((128, 128), (256, 128), (64, 128), (128, 64), (128, 32), (180, 64), (192, 128)):
        for depth in (1, 2, 4, 8):
            if depth * box_rows * row_bytes > 200 * 1024:
                continue
            iters = 400
            for _ in range(2):
                rc = lib.exp_tma_bw(mat.data_ptr(), rows, 64, box_rows, row_bytes, depth, iters, grid, cyc.data_ptr())
                assert rc == 0, rc
            c = cyc[:grid].max().item()
            print(f"{grid:4d} {box_rows:4d} {row_bytes:4d} {depth:2d} | {c / iters:9.1f} {box_rows * row_bytes * iters / c:8.2f}", flush=True)
