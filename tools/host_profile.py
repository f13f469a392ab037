"""cProfile of the python shim around dmc_forward (which part of a module call is host time)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dmc_b200 as D  # noqa: E402

dev = torch.device("cuda:0")
H, W = 1280, 1920
frames, masks = D.clips.synthetic_clip(5, 1, 3, H, W)
x = torch.cat([frames, masks], 2).to(dev)
torch.manual_seed(1)
mp = D.build_p_model("performance").eval().to(dev)
with torch.no_grad():
    r = mp(x[:, 1], 40, {"frame": x[:, 0, :3].contiguous(), "feature": None}, after_i=True)
    for _ in range(3):
        r = mp(x[:, 2], 32, r["dpb"], after_i=False)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for i in range(8):
        r = mp(x[:, 1 + i % 2], 32, r["dpb"], after_i=False)
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
