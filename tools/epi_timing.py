"""Phase clocks of one epilogue warp of the chain kernel (needs a library built with
DMC_NVCC_EXTRA=-DDMC_EPI_TIMING):  average cycles per 16-column chunk spent in each phase."""
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
NAMES = ["slot wait+prefetch", "tcgen05.ld+bias", "combine/act", "residual join", "scale/split/sts/fence", "tma store",
         "chunks", "accumulator wait"]
for name, k, n, mode in [("256->256 plain", 256, 256, 0), ("256->256 wsilu", 256, 256, 1), ("256->256 +res", 256, 256, 2),
                         ("256->1024 pair", 256, 1024, 3), ("512->256 +res", 512, 256, 2)]:
    for probe in (0, 3):
        buf = (ctypes.c_ulonglong * 8)()
        lib.dmc_debug_epi_timing(buf, 1)
        ms = ctypes.c_float()
        rc = lib.dmc_bench_gemm(M, k, n, mode, 3, 2, 5, probe, ctypes.byref(ms))
        lib.dmc_debug_epi_timing(buf, 0)
        ch = max(1, buf[6])
        print(f"{name:16s} probe={probe} {ms.value * 1e3:6.1f} us/launch  chunks={ch}  " +
              "  ".join(f"{NAMES[i]}={buf[i] / ch:6.0f}" for i in (7, 0, 1, 2, 3, 4, 5)), flush=True)
