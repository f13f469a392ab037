"""One training step (forward in train mode + loss + backward) of the reference's `performance` P-frame model at full size:
the stock modules (oracle/_ref, unmodified) under torch eager -- fp32 with TF32 off (the parity-grade arithmetic), default
TF32, autocast bf16 -- against the same class with the engine's training blocks swapped in
(`dmc_b200.training.reference_patched` + `adopt`).  Measurement tool: CUDA events around whole steps, the loss is the
trainer's (trainer_seg_video_model.py:904-934).  One JSON line per configuration.

    tests/diag/train_step_bench.py [H] [W] [iters]      (under tests/: it imports oracle/_ref, the checker)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import dmc_b200 as D  # noqa: E402
from oracle import make_ref  # noqa: E402  (the unmodified reference modules: the thing being compared against)

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
ITERS = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
T = D.training
R = make_ref.load_reference()

frames, masks = D.clips.synthetic_clip(2, 1, 3, H, W)
x = torch.cat([frames, masks], 2).to(dev)
target = x[:, 1, :3].contiguous()
torch.manual_seed(3)
dpb = {"frame": x[:, 0, :3].contiguous(), "feature": torch.randn(1, 256, H // 8, W // 8, device=dev) * 0.5}


def step(model, amp=None):
    model.zero_grad(set_to_none=True)
    if amp:
        with torch.autocast("cuda", dtype=amp):
            r = model(x[:, 1], 32, dpb, after_i=False)
    else:
        r = model(x[:, 1], 32, dpb, after_i=False)
    loss = r["bpp_y"].mean() + r["bpp_z"].mean() + 256.0 * F.mse_loss(r["dpb"]["frame"].float(), target)
    loss.backward()
    return loss


def time_steps(model, amp=None):
    step(model, amp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(ITERS):
        loss = step(model, amp)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / ITERS, float(loss)


torch.manual_seed(11)
stock = R["performance"]().to(dev).train()
mods = [sys.modules[n] for n in ("src.layers.layers", "src.refactor.common_model", "src.refactor.seg_video_model")]
with T.reference_patched(*mods):
    ours = R["performance"]().to(dev).train()
T.adopt(ours, formula=1)
ours.load_state_dict(stock.state_dict())

lib = D._capi.load()
only = os.environ.get("DMC_TS_ONLY", "")
if only in ("", "engine"):
    l0 = lib.dmc_kernel_launches()
    step(ours)
    launches = lib.dmc_kernel_launches() - l0
    # the blocks are fp32-grade either way; what torch still runs between them is timed at both settings
    for rest, tf32 in (("torch remainder fp32, TF32 off", False), ("torch remainder at default TF32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        ms, loss = time_steps(ours)
        print(json.dumps({"impl": "reference model + dmc_b200 training blocks (3-term split fp16, fp32-grade); " + rest,
                          "H": H, "W": W, "train_step_ms": round(ms, 2), "loss": loss,
                          "engine_kernel_launches_first_step": int(launches),
                          "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
T.release_handles()
del ours
torch.cuda.empty_cache()
if only in ("", "stock"):
    for name, tf32, amp in (("stock reference, torch eager fp32 (TF32 off)", False, None),
                            ("stock reference, torch eager default TF32", True, None),
                            ("stock reference, torch eager autocast bf16", True, torch.bfloat16)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.reset_peak_memory_stats()
        try:
            ms, loss = time_steps(stock, amp)
            print(json.dumps({"impl": name, "H": H, "W": W, "train_step_ms": round(ms, 2), "loss": loss,
                              "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
        except Exception as ex:          # (autocast can trip the reference's own NaN guards)
            print(json.dumps({"impl": name, "error": str(ex)[:200]}), flush=True)
