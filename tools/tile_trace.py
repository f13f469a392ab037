"""Tile timeline of cluster 0 of the last chain launch (library built with DMC_NVCC_EXTRA=-DDMC_EPI_TIMING):
when the MMA warp got its TMEM buffer / first operands / issued its last MMA, and when epilogue warp 0 got the
accumulator and finished.   tile_trace.py [H W C]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import dmc_b200 as D  # noqa: E402

lib = D._capi.load()
torch.zeros(1, device="cuda")
h, w, c = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (160, 240, 256)))
ms = ctypes.c_float()
rc = lib.dmc_bench_dcb(1, h, w, c, c, int(os.environ.get("TRACE_BLOCKS", "2")), 3, ctypes.byref(ms))
print("rc", rc, "us/block", ms.value * 1e3)
N = 64
buf = (ctypes.c_ulonglong * (14 * N))()
lib.dmc_debug_tile_trace(buf, N)
t0 = buf[0]
print("tile layer nt  mt | mma: wait_tmem  got_tmem  first_ops  last_issue | epi: wait_acc  got_acc  done | mma_dur epi_dur period")
prev_done = None
for t in range(N):
    r = [buf[14 * t + i] for i in range(14)]
    if r[0] == 0 and t > 0:
        break
    e = r[7]
    rel = [int(x) - int(t0) for x in r[:7]]
    period = "" if prev_done is None else rel[6] - prev_done
    prev_done = rel[6]
    print(f"{t:4d} {e >> 28:5d} {(e >> 20) & 0xff:2d} {e & 0xfffff:4d} | {rel[0]:9d} {rel[1]:9d} {rel[2]:9d} {rel[3]:9d} | "
          f"{rel[4]:9d} {rel[5]:9d} {rel[6]:9d} | {rel[3] - rel[1]:6d} {rel[6] - rel[5]:6d} {period}  | opwait {r[8]:6d} issue {r[9]:6d} | producer: wait-empty {r[11]:6d} issue {r[12]:6d} | last TMEM release of CTA0 {int(r[13]) - int(t0):9d}")

wb = (ctypes.c_ulonglong * (2 * 2 * 256 * 3))()
lib.dmc_debug_warp_trace(wb)
print("tile | CTA0 warp0: got rel done | CTA0 warp7 | CTA1 warp0 | CTA1 warp7   (clocks rel. to t0; CTA1 clock is its own SM's)")
for t in range(0, 40):
    row = []
    for cta in range(2):
        for wi in range(2):
            base = ((cta * 2 + wi) * 256 + t) * 3
            ref = t0 if cta == 0 else wb[((2 + 0) * 256 + 0) * 3] - (buf[5] - t0)
            row.append(" ".join(f"{int(wb[base + i]) - int(ref):7d}" for i in range(3)))
    print(f"{t:4d} | " + " | ".join(row))
