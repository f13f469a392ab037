"""One steady-state 1920x1280 `performance` P-frame inside a cudaProfilerStart/Stop window
(run under `ncu --profile-from-start off ...`; plain runs just print the frame time)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dmc_b200 as D  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "performance"
B, H, W = 1, 1280, 1920
dev = torch.device("cuda:0")
frames, masks = D.clips.synthetic_clip(1000, B, 4, H, W)
x = torch.cat([frames, masks], 2).to(dev) if variant != "old" else frames.to(dev)
u8 = torch.randint(0, 256, (1, H, W, 3), dtype=torch.uint8, device=dev)
m8 = torch.randint(0, 2, (1, H, W), dtype=torch.uint8, device=dev)
torch.manual_seed(1)
mp = D.build_p_model(variant).eval().to(dev)
with torch.no_grad():
    r = mp(x[:, 1], 40, {"frame": x[:, 0, :3].contiguous(), "feature": None}, after_i=True)
    r = mp(x[:, 2], 32, r["dpb"], after_i=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()
    e0.record()
    r = mp(x[:, 3], 36, r["dpb"], after_i=False)
    e1.record()
    D.data.frames_from_u8(u8, m8)                 # (the dataset-side kernel, for its ncu line)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print(f"{variant} P-frame {e0.elapsed_time(e1):.3f} ms, bpp {r['bpp'].tolist()}")
