"""Epilogue of the chain kernel in isolation (no operand loads, no MMAs): all of it, without its TMA traffic (arithmetic
and st.shared only), and the accumulator hand-over alone -- which part of an epilogue-bound tile is the store path?"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import dmc_b200 as D  # noqa: E402

lib = D._capi.load()
torch.zeros(1, device="cuda")
M = 38400
for name, k, n, mode in [("256->256 plain", 256, 256, 0), ("256->256 wsilu", 256, 256, 1), ("256->256 +res", 256, 256, 2),
                         ("256->1024 pair", 256, 1024, 3)]:
    row = []
    for probe, pname in [(3, "epilogue only"), (7, "... without TMA loads/stores"), (11, "hand-over only")]:
        ms = ctypes.c_float()
        rc = lib.dmc_bench_gemm(M, k, n, mode, 3, 2, 20, probe, ctypes.byref(ms))
        row.append(f"{pname}={ms.value * 1e3:6.1f}us" if rc == 0 else f"{pname}=ERR")
    print(f"{name:16s} " + "  ".join(row), flush=True)
