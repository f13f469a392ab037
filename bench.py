#!/usr/bin/env python
"""Headline benchmark: P-frames/s (encoder + decoder + bit estimate = DMC.forward) at 1920x1280.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--clips 64]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (the same for every N; BASELINE.json configs[1] in the multi-clip form of configs[4]):
64 independent `dmc_variant=performance` clips of synthetic Waymo-shaped 1920x1280 frames + synthetic masks,
random-init weights, batch 1 per forward, clip c on rank c mod N (strong scaling: the job is fixed, a rank owns
64/N clips; weights replicated, no data-path collective).  A STEP advances every clip by one GOP position = 64
`forward` calls over the whole job.  Every clip runs real GOPs of 32 frames (position 0: the intra model DMCI,
position 1: the P frame with after_i=True, positions 2..31: P frames); the warm-up runs the GOP head, the timed
region starts at position 2, so a default run (K + 2 <= 32) times P frames only, and a longer one (--steps 64)
contains the I frames of the following GOPs (their time counts, their frames do not: the metric is P-frames/s).

`value`  inputs resident in HBM; includes the caller-side statistics kernel per frame and ONE all-reduce of the
         7-double statistics vector (NCCL) at the end of the timed region.
`e2e`    the same loop through the public nn.Module API with every frame + mask coming from pinned HOST memory
         (H2D inside the timed region, overlapped with the previous forward) and the bpp read back to the host.
         x_hat / feature stay on the device: that is the trainer's contract (it detaches and re-feeds the dpb,
         trainer_seg_video_model.py:1165) -- no image is returned to the host by the reference loop either.

`--impl reference` times the reference's own CPU implementation of the same forward on rank 0: the UNMODIFIED
reference modules from oracle/_ref (copied there by oracle/make_ref.py; kind "reference"), or the oracle port when
that copy is absent (kind "port"), all host threads, each step one full-size P-frame forward.
N = 1 also reports `gpu_eager_baseline`: the same unmodified reference modules run eagerly by PyTorch on the same
B200 (fp32 with TF32 off = the parity-grade arithmetic, default TF32, autocast bf16), and the two training-mode figures of
SURVEY 8(f) rank 2 (not part of the headline): `training_block` (one DepthConvBlock forward + backward, engine against
the layer's arithmetic in torch eager) and `training_step` (forward in train mode + loss + backward of the reference's own
`performance` model class at full size, built from the engine's training blocks against the stock class).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")

import torch  # noqa: E402

H, W, B = 1280, 1920, 1
VARIANT = "performance"
BASE_QP = 32
GOP = 32
DATA_FRAMES = 4            # distinct frames kept per clip (position p uses frame p % 4; position 0 frame 0)
MAX_DISTINCT = 8           # distinct synthetic clips generated per rank (clip slot j shows content j % 8)
METRIC = "P-frames/sec enc+dec @1920x1280"
# 2*MAC over every conv2d of one `performance` P-frame (after_i=False)
ALGO_GFLOP_PER_FRAME = 1092.1


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p.get("bf16_tflops_sustained", p.get("bf16_tflops")), "hbm_gbs": p.get("hbm_gbs"),
                "source": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"bf16_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def ncu_traffic():
    """dram bytes (read + write) per launch of the dominant kernel from the newest committed `ncu --set full` capture."""
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "chain_dcb_r*_ncu_summary.json")))
    for path in reversed(cands):
        try:
            d = json.load(open(path))
            return {"bytes": d["dram_bytes_read"] + d["dram_bytes_write"], "launch": d["launch"],
                    "algorithmic_bytes": d["algorithmic_bytes"], "source": "profiles/" + os.path.basename(path)}
        except Exception:   # noqa: BLE001
            continue
    return None


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.power, self.power_limit = [], None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.h) / 1000.0
            except Exception:   # noqa: BLE001
                pass
        except Exception:   # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                except Exception:   # noqa: BLE001
                    pass
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:   # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        out = {"sm_mhz": statistics.median(self.samples), "sm_mhz_min": min(self.samples),
               "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        if self.power:
            # (board power against its enforced limit: with sw_power_cap set the frame rate is bound by energy per frame)
            out["power_w"] = round(statistics.median(self.power), 1)
            out["power_limit_w"] = self.power_limit
        return out


# ----------------------------------------------------------------------------------------------------------
# reference arms
# ----------------------------------------------------------------------------------------------------------
def _reference_p_model(device="cpu"):
    """(forward(x, qp, dpb, after_i), kind): the unmodified reference module from oracle/_ref, or the oracle port."""
    import dmc_b200 as D
    from oracle import make_ref
    if make_ref.available():
        R = make_ref.load_reference()
        torch.manual_seed(1)
        model = R[VARIANT]().eval().to(device)
        return (lambda x, qp, dpb, after_i: model(x, qp, dpb, after_i=after_i)), "reference"
    from oracle import dmc_oracle as O
    torch.manual_seed(1)
    sd = {k: v.detach().to(device) for k, v in D.build_p_model(VARIANT).state_dict().items()}
    return (lambda x, qp, dpb, after_i: O.dmc_forward(sd, VARIANT, x, qp, dpb, after_i=after_i)), "port"


def cpu_reference_fps(steps, warmup, budget_s=240.0):
    """The reference's CPU path: full-size P-frame forwards (after_i=False), all host threads.  `steps` is cut to
    what fits `budget_s` (the sample that ran is reported)."""
    import dmc_b200 as D
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fwd, kind = _reference_p_model("cpu")
    frames, masks = D.clips.synthetic_clip(1000, B, 3, H, W)
    with torch.no_grad():
        x1 = torch.cat([frames[:, 1], masks[:, 1]], 1)
        x2 = torch.cat([frames[:, 2], masks[:, 2]], 1)
        t0 = time.perf_counter()
        r = fwd(x1, BASE_QP + 8, {"frame": frames[:, 0], "feature": None}, True)        # GOP head, untimed
        t_head = time.perf_counter() - t0
        dpb = {k: (v.detach() if v is not None else None) for k, v in r["dpb"].items()}
        n_warm = min(max(0, warmup), 1 if t_head * (warmup + steps) > budget_s else warmup)
        for _ in range(n_warm):
            fwd(x2, BASE_QP, dpb, False)
        n = max(1, min(steps, int((budget_s - t_head * (1 + n_warm)) / max(t_head, 1e-3))))
        t0 = time.perf_counter()
        for i in range(n):
            fwd(x2 if i % 2 == 0 else x1, BASE_QP, dpb, False)
        t = (time.perf_counter() - t0) / n
    what = "unmodified reference modules (oracle/_ref)" if kind == "reference" else "oracle port (oracle/dmc_oracle.py)"
    sample = (f"{n} full-size {W}x{H} `{VARIANT}` P-frame forwards (after_i=False, batch 1) after {n_warm} warm-up; "
              f"{what}, torch {torch.__version__} fp32, {threads} threads")
    return B / t, threads, sample, t * 1e3, kind


def training_block(dev, C=256, H=160, W=240, iters=20):
    """SURVEY 8(f) rank 2 beside the headline: forward + backward of one DepthConvBlock of the frame (256 channels at
    H/8 x W/8) through dmc_b200.training, and the layer's arithmetic (layers.py:43-79) in stock torch ops under
    autograd on the same GPU.  CUDA events around `iters` whole passes after 3 warm-up passes; every pass streams > 1 GB."""
    import torch.nn.functional as F
    from torch import nn
    import dmc_b200 as D

    def wsilu(v):
        return F.silu(4.0 * v) / 4.0

    class TorchDCB(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.dc0, self.dc2, self.dc3 = nn.Conv2d(c, c, 1), nn.Conv2d(c, c, 3, padding=1, groups=c), nn.Conv2d(c, c, 1)
            self.ffn0, self.ffn2 = nn.Conv2d(c, 4 * c, 1), nn.Conv2d(2 * c, c, 1)

        def forward(self, x, qs):
            out = self.dc3(self.dc2(wsilu(self.dc0(x)))) + x
            u1, u2 = torch.chunk(wsilu(self.ffn0(out)), 2, dim=1)
            return (self.ffn2(u1 + u2) + out) * qs

    x = torch.randn(1, C, H, W, device=dev)
    qs = torch.rand(1, C, 1, 1, device=dev) + 0.5
    gout = torch.randn(1, C, H, W, device=dev) * 1e-6

    def timed(block, amp=None):
        xin, q = x.clone().requires_grad_(True), qs.clone().requires_grad_(True)

        def step():
            xin.grad = None
            for p in block.parameters():
                p.grad = None
            if amp is not None:
                with torch.autocast("cuda", dtype=amp):
                    y = block(xin, q)
            else:
                y = block(xin, q)
            y.backward(gout)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            step()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    out = {"what": "DepthConvBlock(%d) forward + backward at %dx%d, B = 1, ms per pass" % (C, H, W),
           "algorithmic_gflop_per_pass": 3 * 2.0 * H * W * C * C * 8 / 1e9}
    lib = D._capi.load()
    l0 = lib.dmc_kernel_launches()
    blk = D.training.DepthConvBlock(C, C).to(dev).train()
    out["dmc_b200_ms"] = timed(blk)
    out["gpu_launches"] = int(lib.dmc_kernel_launches() - l0)
    del blk
    D.training.release_handles()
    ref = TorchDCB(C).to(dev).train()
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for key, tf32, amp in (("torch_fp32_tf32_off_ms", False, None), ("torch_default_tf32_ms", True, None),
                           ("torch_autocast_bf16_ms", True, torch.bfloat16)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        out[key] = timed(ref, amp)
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    return out


def gpu_eager_baseline(dev, x_frames, n=5):
    """The unmodified reference modules, torch-eager on this GPU, same clip: P-frame forwards (after_i=False)."""
    from oracle import make_ref
    if not make_ref.available():
        return {"unavailable": "oracle/_ref not present (python oracle/make_ref.py where /root/reference exists)"}
    out = {"source": "oracle/_ref: unmodified reference nn.Modules, torch %s eager, CUDA events, %d timed P-frame "
                     "forwards (after_i=False) after 2 warm-up; same clip and weights seed" % (torch.__version__, n),
           "unit": "frames/s"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
    try:
        fwd, _ = _reference_p_model(dev)
        modes = (("fp32_tf32_off", False, False, "highest", False),
                 ("fp32_tf32_default", True, True, "medium", False),     # trainer_seg_video_model.py:59
                 ("autocast_bf16", True, True, "medium", True))
        for name, cudnn_tf32, mm_tf32, prec, amp in modes:
            torch.backends.cudnn.allow_tf32 = cudnn_tf32
            torch.backends.cuda.matmul.allow_tf32 = mm_tf32
            torch.set_float32_matmul_precision(prec)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                r = fwd(x_frames[:, 1], BASE_QP + 8, {"frame": x_frames[:, 0, :3].contiguous(), "feature": None}, True)
                dpb = r["dpb"]
                for _ in range(2):
                    dpb = fwd(x_frames[:, 2], BASE_QP, dpb, False)["dpb"]
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(n):
                    r = fwd(x_frames[:, 1 + i % 3], BASE_QP, dpb, False)
                    dpb = r["dpb"]
                e1.record()
                torch.cuda.synchronize(dev)
            out[name] = n * B / (e0.elapsed_time(e1) / 1e3)
            out[name + "_bpp"] = float(r["bpp"].float().mean())
        del fwd
    except Exception as ex:   # noqa: BLE001
        out["error"] = repr(ex)[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old[0], old[1]
        torch.set_float32_matmul_precision(old[2])
        torch.cuda.empty_cache()
    return out


def training_step(dev, x_frames, n=3):
    """SURVEY 8(f) rank 2 at model scale: one training step (forward in train mode + the trainer's loss + backward) of the
    reference's OWN `performance` model class at full size -- built from dmc_b200.training's blocks and convolutions
    (training.reference_patched + adopt) against the stock class (oracle/_ref, unmodified) in torch eager.  ms per step,
    CUDA events around `n` steps after two warm-up steps."""
    import torch.nn.functional as F
    import dmc_b200 as D
    from oracle import make_ref
    if not make_ref.available():
        return {"unavailable": "oracle/_ref not present"}
    out = {"what": "performance P-frame model, 1920x1280, B = 1, after_i = False: forward (train mode) + bpp_y + bpp_z + "
                   "256 * mse + backward; ms per step"}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.get_float32_matmul_precision())
    T = D.training
    try:
        R = make_ref.load_reference()
        x = x_frames[:, 1]
        target = x[:, :3].contiguous()
        torch.manual_seed(3)
        dpb = {"frame": x_frames[:, 0, :3].contiguous(),
               "feature": torch.randn(x.shape[0], 256, x.shape[2] // 8, x.shape[3] // 8, device=dev) * 0.5}
        torch.manual_seed(11)
        stock = R["performance"]().to(dev).train()
        mods = [sys.modules[k] for k in ("src.layers.layers", "src.refactor.common_model", "src.refactor.seg_video_model")]
        with T.reference_patched(*mods):
            ours = R["performance"]().to(dev).train()
        T.adopt(ours, formula=1)
        ours.load_state_dict(stock.state_dict())

        def timed(model, amp):
            def step():
                model.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                    r = model(x, BASE_QP, dpb, after_i=False)
                loss = r["bpp_y"].mean() + r["bpp_z"].mean() + 256.0 * F.mse_loss(r["dpb"]["frame"].float(), target)
                loss.backward()
                return loss
            for _ in range(2):         # (the engine captures its backward graphs in the second step)
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                loss = step()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) / n, float(loss.detach())

        lib = D._capi.load()
        # the engine's blocks and convolutions are fp32-grade either way; the few ops torch still runs between them are
        # timed once at the parity-grade setting (TF32 off) and once at the trainer's default (TF32)
        l0 = lib.dmc_kernel_launches()
        for name, tf32, prec in (("dmc_b200_blocks_rest_fp32_tf32_off", False, "highest"),
                                 ("dmc_b200_blocks_rest_tf32_default", True, "medium")):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.set_float32_matmul_precision(prec)
            out[name + "_ms"], out[name + "_loss"] = timed(ours, False)
        out["gpu_launches_per_step"] = int((lib.dmc_kernel_launches() - l0) // (2 * (n + 2)))
        T.release_handles()
        del ours
        torch.cuda.empty_cache()
        for name, tf32, prec, amp in (("stock_fp32_tf32_off", False, "highest", False),
                                      ("stock_fp32_tf32_default", True, "medium", False),
                                      ("stock_autocast_bf16", True, "medium", True)):
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.set_float32_matmul_precision(prec)
            out[name + "_ms"], out[name + "_loss"] = timed(stock, amp)
        del stock
    except Exception as ex:   # noqa: BLE001
        out["error"] = repr(ex)[:300]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old[0], old[1]
        torch.set_float32_matmul_precision(old[2])
        T.release_handles()
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clips", type=int, default=64, help="independent clips of the whole job (BASELINE config 5: 64)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--no-training-block", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed resident region (ncu --profile-from-start off)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_clips = max(world, args.clips)
    config = {"workload": f"configs[1] as the clip sweep of configs[4]: dmc_variant={VARIANT}, {W}x{H}, batch {B} per "
                          f"forward, {n_clips} independent clips (clip c on rank c mod N), one step = one GOP position "
                          f"of every clip = {n_clips} forward calls over the job; GOP {GOP} (position 0 DMCI intra, 1 "
                          f"P after_i=True, 2.. P after_i=False, qp {BASE_QP}+shift), timed region starts at position 2; "
                          f"synthetic clips + masks, random-init weights",
              "l2": "per-forward working set (~4 GB of activations) >> 126 MB L2, and consecutive forwards belong to "
                    "different clips: no explicit flush",
              "parallelism": f"{n_clips} clips sharded over {world} GPU(s) ({n_clips // world}-{-(-n_clips // world)} per "
                             f"rank), weights replicated, one all-reduce of a 7-double statistics vector"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        fps, threads, sample, ms, kind = cpu_reference_fps(args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import ctypes
    import torch.distributed as dist
    import dmc_b200 as D
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: a CUDA device is required (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = D._capi.load()

    torch.manual_seed(0)
    mi = D.DMCI().eval().to(dev)
    torch.manual_seed(1)
    mp = D.build_p_model(VARIANT).eval().to(dev)
    my_clips = D.clips.shard_clips(n_clips, rank, world)
    n_local = len(my_clips)
    n_distinct = min(n_local, MAX_DISTINCT)
    # (T,4,H,W) per distinct clip: pinned host copy (e2e) and device copy (resident run)
    host, devc = [], []
    for j in range(n_distinct):
        f, m = D.clips.synthetic_clip(1000 + my_clips[j], B, DATA_FRAMES, H, W)
        x = torch.cat([f, m], dim=2).contiguous().pin_memory()
        host.append(x)
        devc.append(x.to(dev))
    stats = D.clips.ClipStats(dev)
    dpbs = [None] * n_local

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def frame_of(j, pos):
        return devc[j % n_distinct][:, pos % DATA_FRAMES if pos else 0]

    def forward(j, pos, x):
        """GOP position `pos` of clip slot j on input x (B,4,H,W).  Returns (result, number of P frames)."""
        if pos == 0:
            r = mi(x[:, :3].contiguous(), BASE_QP)
            dpbs[j] = r["dpb"]
            return r, 0
        qp = mp.shift_qp(BASE_QP, D.clips.INDEX_MAP[pos % 8])
        r = mp(x, qp, dpbs[j], after_i=(pos == 1))
        dpbs[j] = r["dpb"]
        return r, B

    def run_resident(first_pos, n_steps, with_stats):
        p_frames = 0
        for s in range(n_steps):
            pos = (first_pos + s) % GOP
            for j in range(n_local):
                x = frame_of(j, pos)
                r, n = forward(j, pos, x)
                p_frames += n
                if with_stats and n:
                    stats.add_frame(r, x[:, :3], x[:, 3:4])
        return p_frames

    copy_stream = torch.cuda.Stream(dev)

    def prefetch(j, pos):
        """H2D of the frame + mask of (clip j, position pos) from pinned host memory, on the copy stream."""
        with torch.cuda.stream(copy_stream):
            x = host[j % n_distinct][:, pos % DATA_FRAMES if pos else 0].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return x, ev

    def run_host(first_pos, n_steps):
        """The same loop end to end, software-pipelined two deep: the H2D copy of forward i+1 overlaps the kernels
        of forward i, and the host reads the bpp of forward i (after its D2H) while forward i+1 is already queued."""
        outs = [torch.empty(B, 3).pin_memory() for _ in range(2)]
        work = [(j, (first_pos + s) % GOP) for s in range(n_steps) for j in range(n_local)]
        total, p_frames, prev = 0.0, 0, None
        nxt = prefetch(*work[0])
        for i, (j, pos) in enumerate(work):
            x, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            x.record_stream(torch.cuda.current_stream())
            r, n = forward(j, pos, x)
            p_frames += n
            outs[i & 1].copy_(torch.stack([r["bpp"], r["bpp_y"], r["bpp_z"]], 1), non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            if i + 1 < len(work):
                nxt = prefetch(*work[i + 1])
            if prev is not None:
                prev[0].synchronize()
                total += float(prev[1][0, 0])              # the caller consumes the bpp of every forward
            prev = (done, outs[i & 1])
        prev[0].synchronize()
        total += float(prev[1][0, 0])
        return p_frames, total

    with torch.no_grad():
        # ---- warm-up: W steps starting with the GOP head (I frame, first P frame), extended by whole steps until the
        # device has been busy for ~1 s (a B200 that idled through model construction needs that long to settle its
        # clocks; the first 20 forwards of a fresh process measured 6 % slower than the next 20 otherwise)
        t_warm = time.time()
        n_warm = 0
        while n_warm < max(3, args.warmup) or time.time() - t_warm < 1.0:
            run_resident(n_warm if n_warm < 2 else 2 + (n_warm - 2) % (GOP - 2), 1, False)
            n_warm += 1
            torch.cuda.synchronize(dev)
        # ---- timed: resident inputs
        barrier()
        l0 = lib.dmc_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            if args.profiler_range:
                torch.cuda.profiler.start()
            e0.record()
            p_res = run_resident(2, args.steps, True)
            stats.all_reduce()
            e1.record()
            barrier()
            if args.profiler_range:
                torch.cuda.profiler.stop()
        launches = lib.dmc_kernel_launches() - l0
        ms = e0.elapsed_time(e1)
        mp.check_finite()
        # ---- timed: end to end through the public API with host buffers
        run_host(2, 1)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk2:
            e2.record()
            p_e2e, _ = run_host(2, args.steps)
            e3.record()
            barrier()
        ms_e2e = e2.elapsed_time(e3)
        # ---- contraction kernel: device time per launch (CUDA events on the launching stream)
        h, _ = mp._engine(B, H, W, dev)
        lib.dmc_profile_enable(h, 1)
        nprof = 3
        for i in range(nprof):
            forward(0, 2 + i, frame_of(0, 2 + i))
        g_ms, g_n, g_fl, g_is = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
        lib.dmc_profile_read(h, ctypes.byref(g_ms), ctypes.byref(g_n), ctypes.byref(g_fl), ctypes.byref(g_is))
        lib.dmc_profile_enable(h, 0)

    c1, c2 = clk.summary(), clk2.summary()
    mine = torch.tensor([ms, ms_e2e, float(p_res), float(p_e2e), float(c1["sm_mhz"] or 0), float(c1.get("sm_mhz_min") or 0),
                         float(c2["sm_mhz"] or 0)], dtype=torch.float64, device=dev)
    if world > 1:
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        reasons = [None] * world
        dist.all_gather_object(reasons, sorted(set(c1["reasons"]) | set(c2["reasons"])))
    else:
        allr, reasons = [mine], [sorted(set(c1["reasons"]) | set(c2["reasons"]))]
    allr = torch.stack(allr).cpu()
    ms, ms_e2e = float(allr[:, 0].max()), float(allr[:, 1].max())          # MAX over ranks
    p_total, p_total_e2e = float(allr[:, 2].sum()), float(allr[:, 3].sum())
    value = p_total / (ms / 1e3)
    e2e = p_total_e2e / (ms_e2e / 1e3)
    if rank == 0:
        pk = peaks()
        per_launch_ms = g_ms.value / max(1, g_n.value)
        achieved = (g_fl.value / max(1, g_n.value)) / (per_launch_ms * 1e-3) / 1e12 if g_n.value else 0.0
        summ = stats.summary()
        fwd_per_step = p_total / args.steps
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None,
                "dtype": "f16 tensor-core operands (tcgen05 kind::f16), fp32 accumulate: 3-term split product (fp16 hi + "
                         "2^11-scaled fp16 lo) on every layer that can reach a symbol; single term in recon_generation_net",
                "data": "synthetic", "config": dict(config, warmup_steps_run=n_warm, clips=n_clips,
                                                    forwards_per_step=fwd_per_step,
                                                    ms_per_p_frame=ms / args.steps / (fwd_per_step / world)),
                "clocks": c1,
                "per_rank": [{"rank": r, "ms": float(allr[r, 0]), "ms_e2e": float(allr[r, 1]),
                              "p_frames": int(allr[r, 2]), "sm_mhz": float(allr[r, 4]), "sm_mhz_min": float(allr[r, 5]),
                              "sm_mhz_e2e": float(allr[r, 6]), "reasons": reasons[r]} for r in range(world)],
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(fwd_per_step) * B * 4 * H * W * 4,
                        "d2h_bytes_per_step": int(fwd_per_step) * B * 3 * 4,
                        "note": "per forward: 39.3 MB frame+mask H2D from pinned memory, 12 B bpp D2H; x_hat / feature "
                                "stay on the device as in the reference's GOP loop (the dpb is re-fed, trainer:1165)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor",
                             "kernel": "k_gemm_s3_chain (persistent tcgen05 chain of 1x1 layers, 3-term split-fp16 product)",
                             "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": achieved / pk["bf16_tflops"] if pk["bf16_tflops"] else None,
                             "traffic": (ncu_traffic() or {}).get("bytes"),
                             "traffic_detail": ncu_traffic(),
                             "peak_source": pk["source"],
                             "launches_per_frame": g_n.value / nprof, "avg_launch_ms": per_launch_ms,
                             "gemm_share_of_frame": (g_ms.value / nprof) / (ms / max(1.0, float(allr[0, 2]))),
                             "issued_mma_tflops": g_is.value / (g_ms.value * 1e-3) / 1e12 if g_ms.value else 0.0,
                             "issued_frac": (g_is.value / (g_ms.value * 1e-3) / 1e12 / pk["bf16_tflops"])
                             if g_ms.value and pk["bf16_tflops"] else None,
                             "note": "achieved = algorithmic conv FLOPs (2*M*N*K) per contraction launch / mean launch "
                                     "time (CUDA events around every launch, 3 frames of rank 0); a launch is a chain "
                                     "of 1-5 layers.  fp32-grade layers issue 3 fp16 MMA terms per product, so frac <= "
                                     "1/3 there: issued_mma_tflops / issued_frac count the MMAs actually issued"},
                "quality": {"bpp": summ["bpp"], "psnr": summ["psnr"], "roi_psnr": summ["roi_psnr"],
                            "frames": summ["frames"]},
                "algorithmic_tflops": ALGO_GFLOP_PER_FRAME * value / 1e3}
        if world == 1 and not args.no_eager_baseline:
            del dpbs[:]
            mi.release_engines()
            torch.cuda.empty_cache()
            line["gpu_eager_baseline"] = gpu_eager_baseline(dev, devc[0])
        if world == 1 and not args.no_training_block:
            line["training_block"] = training_block(dev)
            line["training_step"] = training_step(dev, devc[0])
        if world == 1 and not args.no_cpu_baseline:
            fps, threads, sample, _, kind = cpu_reference_fps(3, 0, budget_s=25.0)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
