"""Times the DMCI intra forward at 1920x1280 (CUDA events, steady state)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dmc_b200 as D  # noqa: E402

dev = torch.device("cuda:0")
frames, _ = D.clips.synthetic_clip(5, 1, 2, 1280, 1920)
x = frames[:, 0].to(dev)
torch.manual_seed(0)
mi = D.DMCI().eval().to(dev)
with torch.no_grad():
    for _ in range(3):
        r = mi(x, 32)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = D._capi.load().dmc_kernel_launches()
    e0.record()
    n = 10
    for _ in range(n):
        r = mi(x, 32)
    e1.record()
    torch.cuda.synchronize()
    l1 = D._capi.load().dmc_kernel_launches()
print(f"DMCI 1920x1280: {e0.elapsed_time(e1) / n:.3f} ms per intra frame, {(l1 - l0) // n} launches, bpp {float(r['bpp']):.4f} "
      f"(2390.6 GFLOP algorithmic -> {2390.6 / (e0.elapsed_time(e1) / n):.0f} TFLOP/s)")
