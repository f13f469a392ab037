"""Measures the accumulation error of a single 1x1 contraction against fp64 (not a test):
signed bias (mean of err * sign(ref)) and rms, in units of 2^-24 * |ref| ("ulp"), for the tcgen05
path and the fp32-FMA path, over K and over operand statistics."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import test_gpu_parity as T  # noqa: E402

kappas = [float(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "0,0.276").split(",")]
lib = T.capi.load()
for K in (64, 256, 512, 1024, 2304):
    for kind in ("randn", "positive", "wsilu"):
        g = torch.Generator().manual_seed(K)
        x = torch.randn(1, K, 64, 96, generator=g)
        w = torch.randn(256, K, 1, 1, generator=g) / K ** 0.5
        if kind == "positive":
            x = x.abs()
            w = w.abs()
        if kind == "wsilu":                 # what a layer behind a WSiLU sees: activations mostly positive, weights zero-mean
            x = T.O.wsilu(x)
        ref = F.conv2d(x.double(), w.double())
        big = ref.abs() > 0.25 * float(ref.abs().mean())
        for name, be, kappa in [("tcgen05", 0, k) for k in kappas] + [("simt", 1, 0.0)]:
            lib.dmc_set_acc_comp(kappa)
            name = f"{name} k={kappa:.3f}" if be == 0 else name
            out = T.op_conv2d(x, w, None, backend=be).double()
            rel = ((out - ref) * ref.sign() / ref.abs().clamp_min(1e-30))[big] * 2.0 ** 24
            print(f"K={K:5d} {kind:8s} {name:16s} bias {float(rel.mean()):+8.3f} ulp  rms {float(rel.pow(2).mean().sqrt()):7.3f} ulp  "
                  f"max {float(rel.abs().max()):8.2f}", flush=True)
