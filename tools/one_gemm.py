"""Runs one contraction shape a few times (for ncu captures):  one_gemm.py K N MODE [KERNEL] [ROWS] [PROBE]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import dmc_b200 as D  # noqa: E402

lib = D._capi.load()
torch.zeros(1, device="cuda")
k, n, mode = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 2
rows = int(sys.argv[5]) if len(sys.argv) > 5 else 38400
probe = int(sys.argv[6]) if len(sys.argv) > 6 else 0
ms = ctypes.c_float()
rc = lib.dmc_bench_gemm(rows, k, n, mode, 3, kernel, 5, probe, ctypes.byref(ms))
print("rc", rc, "us/launch", ms.value * 1e3, lib.dmc_last_error(None).decode() if rc else "")
