"""Forward + backward of one DepthConvBlock at frame scale: the engine's training block against the
reference's layer code under torch.autograd on the same GPU (fp32 with TF32 off -- the parity-grade arithmetic --,
default TF32, autocast bf16).  Events on the current stream, L2 flushed between iterations by the working set itself
(> 1 GB per pass).  Prints one JSON line per configuration.      train_block_bench.py [C] [H] [W] [iters]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch import nn  # noqa: E402

import dmc_b200 as D  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 160
W = int(sys.argv[3]) if len(sys.argv) > 3 else 240
ITERS = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda:0")


def wsilu(x):
    return F.silu(4.0 * x) / 4.0


class TorchDCB(nn.Module):          # the arithmetic of src/layers/layers.py:43-79 in stock torch ops
    def __init__(self, c):
        super().__init__()
        self.dc0, self.dc2, self.dc3 = nn.Conv2d(c, c, 1), nn.Conv2d(c, c, 3, padding=1, groups=c), nn.Conv2d(c, c, 1)
        self.ffn0, self.ffn2 = nn.Conv2d(c, 4 * c, 1), nn.Conv2d(2 * c, c, 1)

    def forward(self, x, qs):
        out = self.dc3(self.dc2(wsilu(self.dc0(x)))) + x
        u1, u2 = torch.chunk(wsilu(self.ffn0(out)), 2, dim=1)
        return (self.ffn2(u1 + u2) + out) * qs


def time_it(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


x = torch.randn(1, C, H, W, device=dev)
qs = torch.rand(1, C, 1, 1, device=dev) + 0.5
gout = torch.randn(1, C, H, W, device=dev) * 1e-6
M = H * W
flops_fwd = 2.0 * M * C * C * 8          # dc.0, dc.3: C*C each; ffn.0: 4 C*C; ffn.2: 2 C*C
lib = D._capi.load()


def run(block, xin, q, amp=None):
    def step():
        xin.grad = None
        for p in block.parameters():
            p.grad = None
        if amp:
            with torch.autocast("cuda", dtype=amp):
                y = block(xin, q)
        else:
            y = block(xin, q)
        y.backward(gout)
    return step


def fwd_only(block, xin, q):
    def step():
        with torch.no_grad():
            block(xin, q)
    return step


ONLY = os.environ.get("DMC_TB_ONLY", "")      # "dmc3" / "dmc1": just that engine configuration (profiling runs)
for terms in (3, 1):
    if ONLY and ONLY != f"dmc{terms}":
        continue
    blk = D.training.DepthConvBlock(C, C, terms=terms).to(dev).train()
    xin = x.clone().requires_grad_(True)
    q = qs.clone().requires_grad_(True)
    l0 = lib.dmc_kernel_launches()
    run(blk, xin, q)()
    launches = lib.dmc_kernel_launches() - l0
    ms = time_it(run(blk, xin, q), ITERS)
    ms_f = time_it(fwd_only(blk, xin, q), ITERS)
    print(json.dumps({"impl": "dmc_b200", "terms": terms, "C": C, "H": H, "W": W, "fwd_bwd_ms": round(ms, 4),
                      "fwd_ms": round(ms_f, 4), "kernel_launches_fwd_bwd": int(launches),
                      "algorithmic_tflops_fwd_bwd": round(3 * flops_fwd / ms / 1e9, 1)}), flush=True)
    del blk
    D.training.release_handles()

ref = TorchDCB(C).to(dev).train()
for name, tf32, amp in () if ONLY else (("torch fp32 (TF32 off)", False, None), ("torch default TF32", True, None),
                        ("torch autocast bf16", True, torch.bfloat16)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    xin = x.clone().requires_grad_(True)
    q = qs.clone().requires_grad_(True)
    ms = time_it(run(ref, xin, q, amp), ITERS)
    ms_f = time_it(fwd_only(ref, xin, q), ITERS)
    print(json.dumps({"impl": name, "C": C, "H": H, "W": W, "fwd_bwd_ms": round(ms, 4), "fwd_ms": round(ms_f, 4),
                      "algorithmic_tflops_fwd_bwd": round(3 * flops_fwd / ms / 1e9, 1)}), flush=True)
