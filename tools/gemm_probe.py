"""Times single contraction launches in isolation and with parts of the kernel switched off
(no TMA / no MMA / no epilogue traffic) to see which of the three pipelines bounds it."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import dmc_b200 as D  # noqa: E402

lib = D._capi.load()
torch.zeros(1, device="cuda")
M = int(os.environ.get("PROBE_M", "38400"))
SHAPES = [("dc0 256->256 plain", 256, 256, 0), ("dc0 256->256 wsilu", 256, 256, 1), ("dc3 256->256 +res", 256, 256, 2),
          ("ffn0 256->1024 pair", 256, 1024, 3), ("ffn2 512->256 +res", 512, 256, 2), ("head 320->192", 320, 192, 0)]
PROBES = [(0, "full"), (4, "no-epi-mem"), (12, "no-epi"), (2, "no-mma"), (3, "epi-only"), (13, "mma-only"),
          (14, "tma-only"), (32, "no-W-loads"), (46, "tma-only,no-W"), (13 + 64, "mma-only x2")]
ARGS = sys.argv[1:]
KERNELS = [int(a) for a in ARGS if a.isdigit()] or ([] if ARGS else [0, 2])   # 0 general one-CTA, 1 its pair variant, 2 gemm_s3
for name, k, n, mode in SHAPES:
    flops = 2.0 * M * k * n * 6
    for pair in KERNELS:
        row = []
        for probe, pname in PROBES:
            ms = ctypes.c_float()
            rc = lib.dmc_bench_gemm(M, k, n, mode, 3, pair, 20, probe, ctypes.byref(ms))
            if rc != 0:
                row.append(f"{pname}=ERR({lib.dmc_last_error(None).decode()[:60]})")
                break
            row.append(f"{pname}={ms.value * 1e3:7.1f}us")
        print(f"{name:22s} kernel={pair} issued={flops / 1e9:6.1f}GF  " + "  ".join(row), flush=True)

for C in ((256, 320, 128) if ("dw" in ARGS or not ARGS) else ()):
    for f32 in (0, 1):
        ms = ctypes.c_float()
        hh, ww = (160, 240) if C != 128 else (80, 120)
        rc = lib.dmc_bench_dwconv(1, hh, ww, C, f32, 20, ctypes.byref(ms))
        gb = hh * ww * C * (10 if f32 else 12) / 1e9
        print(f"dwconv3x3 {hh}x{ww}x{C} f32_in={f32}: {ms.value * 1e3:7.1f}us  {gb / ms.value * 1e3:7.0f} GB/s (read+write)" if rc == 0
              else f"dwconv ERR {lib.dmc_last_error(None).decode()}", flush=True)

if "dcb" in ARGS or not ARGS:
    for (hh, ww, cin, cout) in ((160, 240, 256, 256), (160, 240, 320, 320), (80, 120, 384, 384)):
        ms = ctypes.c_float()
        rc = lib.dmc_bench_dcb(1, hh, ww, cin, cout, 4, 10, ctypes.byref(ms))
        gf = 2.0 * hh * ww * (2 * cout * cout + 4 * cout * cout + 2 * cout * cout) * 6 / 1e9
        print(f"DepthConvBlock {hh}x{ww} C={cout}: {ms.value * 1e3:7.1f}us per block, issued {gf / ms.value / 1e3:6.0f} TFLOP/s"
              if rc == 0 else f"dcb ERR {lib.dmc_last_error(None).decode()}", flush=True)
