"""Runs BASELINE.json configs 2-5 at full size (1920x1280) on one GPU and prints a JSON line per config:
throughput from CUDA events, sanity properties (finite, ranges) and the clip statistics.

    python tools/run_configs.py [2 3 4 5]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import dmc_b200 as D  # noqa: E402

H, W, QP = 1280, 1920, 32
dev = torch.device("cuda", 0)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1)


def run(cfg):
    variant, batch, frames_n, feedback, clips = {
        2: ("performance", 1, 9, False, 1), 3: ("fast", 8, 5, False, 1),
        4: ("mask_prop", 1, 32, True, 1), 5: ("performance", 1, 5, False, 8)}[cfg]
    torch.manual_seed(0)
    mi = D.DMCI().eval().to(dev)
    torch.manual_seed(1)
    mp = D.build_p_model(variant).eval().to(dev)
    stats = D.clips.ClipStats(dev)
    total_ms, pframes = 0.0, 0
    for clip in range(clips):
        fr, mk = D.clips.synthetic_clip(2000 + 10 * cfg + clip, batch, frames_n, H, W)
        fr, mk = fr.to(dev), mk.to(dev)
        with torch.no_grad():
            i_res = mi(fr[:, 0], QP)
            if clip == 0:        # warm-up: builds the engines
                # (four frames: every qp of the schedule -- a new qp means a new CUDA graph, captured on first use)
                D.clips.run_gop(mi, mp, variant, fr[:, :4], mk[:, :4], QP, None, feedback, i_result=i_res)
            outs, ms = timed(lambda: D.clips.run_gop(mi, mp, variant, fr, mk, QP, stats, feedback, i_result=i_res))
        total_ms += ms
        pframes += (frames_n - 1) * batch
        for o in outs[1:]:
            x = o["dpb"]["frame"]
            assert bool(torch.isfinite(x).all()) and float(x.min()) >= 0 and float(x.max()) <= 1
            assert bool(torch.isfinite(o["bpp"]).all()) and float(o["bpp"].min()) > 0
    s = stats.summary()
    line = {"config": cfg, "variant": variant, "batch": batch, "frames_per_clip": frames_n, "clips": clips,
            "mask_feedback": feedback, "p_frames": pframes, "ms_total": total_ms,
            "p_frames_per_s": pframes / (total_ms * 1e-3), "ms_per_forward": total_ms / ((frames_n - 1) * clips),
            "bpp": s["bpp"], "psnr": s["psnr"], "roi_psnr": s["roi_psnr"]}
    print(json.dumps(line), flush=True)
    del mi, mp
    torch.cuda.empty_cache()


if __name__ == "__main__":
    for c in ([int(a) for a in sys.argv[1:]] or [2, 3, 4, 5]):
        t0 = time.time()
        run(c)
