"""Runs the shifted-descriptor experiment on the GPU box and prints which variants reproduce the
expected row-shifted view."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libexp.so")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "shift_mma.cu")):
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-shared", "-Xcompiler", "-fPIC",
                           "-o", so, os.path.join(HERE, "shift_mma.cu")])
lib = ctypes.CDLL(so)
lib.exp_shift_mma.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4
d = torch.device("cuda")
for sw in (128, 64, 32):
    kc = sw // 2
    X = torch.arange(256 * kc, dtype=torch.float32).reshape(256, kc)
    X = ((X % 2039) / 16.0).to(torch.float16)          # distinct-ish values exactly representable
    Xd = X.to(d)
    Bm = torch.zeros(64, kc, dtype=torch.float16)
    Bm[:kc, :kc] = torch.eye(kc, dtype=torch.float16)
    Bd = Bm.to(d)
    for sbo_rows in (8, 10, 18):
        for mode in (0, 1):
            ok = []
            for shift in range(0, 24):
                D = torch.full((128, 64), -1.0, device=d)
                rc = lib.exp_shift_mma(Xd.data_ptr(), Bd.data_ptr(), D.data_ptr(), shift, sbo_rows * sw, mode, sw)
                if rc:
                    print("rc", rc); sys.exit(1)
                m = torch.arange(128)
                src = (m // 8) * sbo_rows + (m % 8) + shift
                valid = src < 256
                exp = X[src.clamp(max=255)].float()
                got = D.cpu()[:, :kc]
                ok.append(bool(torch.equal(got[valid], exp[valid])))
            print(f"sw={sw} sbo_rows={sbo_rows} base_offset_mode={mode}: shifts ok = {''.join('1' if o else '0' for o in ok)}", flush=True)
