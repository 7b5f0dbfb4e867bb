"""Cost of the pieces of the thread-per-row epilogue (see epilogue_cost.cu): clk per tile per warpgroup for C = 48 / 96,
im2col (128 consecutive pixels) and halo-patch (8 x 16 pixels) row geometry, one or both warpgroups active."""
import ctypes, os, subprocess
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp_epi.so")
import sys
if "--no-build" not in sys.argv or not os.path.exists(so):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC",
                           "-o", so, os.path.join(HERE, "epilogue_cost.cu")])
lib = ctypes.CDLL(so)
lib.exp_epilogue_cost.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_void_p]
d = torch.device("cuda")
grid = 148
LD, LD1, MATH, ST, RES, ONE, PATCH = 1, 2, 4, 8, 16, 32, 64
modes = [("ld 2-per-wait", LD), ("ld one wait", LD | LD1), ("ld+math", LD | MATH), ("math+store", MATH | ST),
         ("ld+math+store", LD | MATH | ST), ("ld1+math+store", LD | LD1 | MATH | ST), ("ld+math+res", LD | MATH | RES),
         ("all (product)", LD | MATH | ST | RES), ("all, one wait", LD | LD1 | MATH | ST | RES)]
for C, Wmap, Hmap in ((48, 72, 96), (64, 72, 96)):   # (the C = 96 instantiation spills: whole-tile register arrays)
    n = 64
    tiles = n * Hmap * Wmap // 128
    res = torch.randn(n * Hmap * Wmap, C, device=d).to(torch.float16)
    out = torch.empty_like(res)
    cyc = torch.zeros(grid * 2 * 2, dtype=torch.int64, device=d)
    print(f"C={C} map {Hmap}x{Wmap}, {tiles} tiles on {grid} CTAs")
    for geo, gname in ((0, "im2col rows"), (PATCH, "patch rows")):
        if geo and (Wmap % 8 or Hmap % 16):
            continue
        for name, m in modes:
            row = []
            for one in (ONE, 0):
                for _ in range(2):
                    cyc.zero_()
                    rc = lib.exp_epilogue_cost(C, res.data_ptr(), out.data_ptr(), tiles, Wmap, m | geo | one, grid, cyc.data_ptr())
                    assert rc == 0, rc
                c = cyc.view(-1, 2).double()
                act = c[:, 1] > 0
                row.append(float((c[act, 0] / c[act, 1]).mean()))
            print(f"  {gname:12s} {name:16s}: one warpgroup {row[0]:7.0f} clk/tile, both {row[1]:7.0f} clk/tile per warpgroup", flush=True)
