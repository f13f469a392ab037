"""Where the time of a full-size training step goes (engine blocks swapped into the reference's `performance` model):
torch.profiler kernel table of one step, grouped into engine kernels and torch / cuDNN kernels.
    tests/diag/train_step_profile.py [H] [W]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import dmc_b200 as D  # noqa: E402
from oracle import make_ref  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
W = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
dev = torch.device("cuda:0")
T = D.training
R = make_ref.load_reference()
frames, masks = D.clips.synthetic_clip(2, 1, 3, H, W)
x = torch.cat([frames, masks], 2).to(dev)
target = x[:, 1, :3].contiguous()
torch.manual_seed(3)
dpb = {"frame": x[:, 0, :3].contiguous(), "feature": torch.randn(1, 256, H // 8, W // 8, device=dev) * 0.5}
mods = [sys.modules[n] for n in ("src.layers.layers", "src.refactor.common_model", "src.refactor.seg_video_model")]
with T.reference_patched(*mods):
    model = R["performance"]().to(dev).train()
T.adopt(model, 1)


def step():
    model.zero_grad(set_to_none=True)
    r = model(x[:, 1], 32, dpb, after_i=False)
    loss = r["bpp_y"].mean() + r["bpp_z"].mean() + 256.0 * F.mse_loss(r["dpb"]["frame"], target)
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count)
        for e in prof.key_averages()]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
eng = sum(r[1] for r in rows if r[0].startswith("dmc::") or "dmc::" in r[0])
print(f"GPU kernel time of one step: {tot / 1000:.2f} ms; engine kernels {eng / 1000:.2f} ms, torch / library kernels "
      f"{(tot - eng) / 1000:.2f} ms")
for k, t, n in rows[:28]:
    print(f"{t / 1000:8.3f} ms  {100 * t / tot:5.1f} %  x{n:<4d} {k[:110]}")
