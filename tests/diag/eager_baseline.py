"""The oracle's torch-fp32 arithmetic (TF32 off) run eagerly on the GPU: the speed a user of the reference gets on
the same B200 without this engine (not a test; the oracle is the checker, this only times it).

    python tests/diag/eager_baseline.py [variant] [frames]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from helpers import D, O, gc, sd_of  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "performance"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
frames, masks = D.clips.synthetic_clip(3, 1, 4, 1280, 1920)
torch.manual_seed(gc.SEED_P)
mp = D.build_p_model(variant).eval()
sd = {k: v.to(dev) for k, v in sd_of(mp).items()}
fr, mk = frames.to(dev), masks.to(dev)
torch.set_default_device(dev)            # the oracle's factory calls (checkerboard masks, constants) follow
for amp in (False, True):
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        dpb = {"frame": fr[:, 0], "feature": None}
        x = lambda t: fr[:, t] if variant == "old" else torch.cat([fr[:, t], mk[:, t]], 1)
        r = O.dmc_forward(sd, variant, x(1), 40, dpb, after_i=True)
        r = O.dmc_forward(sd, variant, x(2), 32, r["dpb"], after_i=False)
        torch.cuda.synchronize()
        t0 = time.time()
        for i in range(n):
            r = O.dmc_forward(sd, variant, x(1 + i % 3), 32, r["dpb"], after_i=False)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / n
    print(f"oracle arithmetic, torch eager on {torch.cuda.get_device_name(0)}, {variant} 1920x1280, "
          f"{'autocast(bf16)' if amp else 'fp32 (TF32 off)'}: {dt * 1e3:.1f} ms/frame = {1 / dt:.2f} P-frames/s, "
          f"bpp {float(r['bpp']):.4f}", flush=True)
