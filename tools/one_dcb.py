"""Runs a stack of DepthConvBlocks a few times (for ncu captures):  one_dcb.py H W C [BLOCKS] [ITERS]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import dmc_b200 as D  # noqa: E402

lib = D._capi.load()
torch.zeros(1, device="cuda")
h, w, c = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
blocks = int(sys.argv[4]) if len(sys.argv) > 4 else 3
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ms = ctypes.c_float()
rc = lib.dmc_bench_dcb(1, h, w, c, c, blocks, iters, ctypes.byref(ms))
print("rc", rc, "us/block", ms.value * 1e3, lib.dmc_last_error(None).decode() if rc else "")
