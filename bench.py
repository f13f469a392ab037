#!/usr/bin/env python
"""Headline benchmark: P-frames/s (encoder + decoder + bit estimate = DMC.forward) at 1920x1280.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one P-frame forward over one batch (B=1) of BASELINE.json configs[1]
(`dmc_variant=performance`, synthetic Waymo-shaped 1920x1280 frames + synthetic masks, random-init
weights).  `value` is measured with the clip resident in HBM; `e2e` is the same step through the
public nn.Module API with the frame+mask in pinned HOST memory (H2D inside the timed region) and
the bpp read back to the host.  Multi-GPU runs shard independent clips across ranks (weak
scaling, no data-path collective) and all-reduce the 7-double statistics vector once.

`--impl reference` times the reference's CPU implementation of the same forward (the oracle port,
oracle/dmc_oracle.py, all host threads) on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
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
CLIP_FRAMES = 6
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
    """dram bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture."""
    path = os.path.join(ROOT, "profiles", "chain_dcb_r01_v8_ncu_summary.json")
    try:
        d = json.load(open(path))
        return {"bytes": d["dram_bytes_read"] + d["dram_bytes_write"], "launch": d["launch"],
                "algorithmic_bytes": d["algorithmic_bytes"], "source": "profiles/" + os.path.basename(path)}
    except Exception:   # noqa: BLE001
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:   # noqa: BLE001
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:   # noqa: BLE001
                pass
            self._stop.wait(0.1)

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
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_fps(steps, warmup, budget_s=150.0):
    """Oracle port (torch fp32, all host threads) on a bounded sample of the same workload."""
    from oracle import dmc_oracle as O
    import dmc_b200 as D
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    sd = {k: v.detach() for k, v in D.build_p_model(VARIANT).state_dict().items()}
    frames, masks = D.clips.synthetic_clip(0, B, 3, H, W)

    def one(h, w, n):
        x = torch.cat([frames[:, 1, :, :h, :w], masks[:, 1, :, :h, :w]], 1)
        dpb = {"frame": frames[:, 0, :, :h, :w], "feature": None}
        r = O.dmc_forward(sd, VARIANT, x, BASE_QP + 8, dpb, after_i=True)
        x2 = torch.cat([frames[:, 2, :, :h, :w], masks[:, 2, :, :h, :w]], 1)
        t0 = time.perf_counter()
        for _ in range(n):
            O.dmc_forward(sd, VARIANT, x2, BASE_QP, r["dpb"], after_i=False)
        return (time.perf_counter() - t0) / n

    # probe on a small crop, then pick the largest crop (sizes multiples of 64) that fits the budget
    crops = [(H, W), (640, 960), (320, 512)]
    t_probe = one(*crops[-1], 1)
    full_area = float(H * W)
    h, w = crops[-1]
    for ch, cw in crops:
        est = t_probe * (ch * cw) / (crops[-1][0] * crops[-1][1])
        if est * (steps + warmup) <= budget_s:
            h, w = ch, cw
            break
    frac = full_area / (h * w)
    if warmup:
        one(h, w, 1)
    t = one(h, w, max(1, steps))
    fps = (1.0 / frac) / t * B
    sample = (f"{max(1, steps)} P-frame forwards (after_i=False) of a {w}x{h} crop = 1/{frac:.2f} of the 1920x1280 "
              f"frame, scaled by area; oracle port, torch {torch.__version__} fp32")
    return fps, threads, sample, t * 1e3 * frac


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed resident region (ncu --profile-from-start off)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"configs[1]: dmc_variant={VARIANT}, {W}x{H}, batch {B}, one P-frame forward per step "
                          f"(after_i=False, qp {BASE_QP}+shift), synthetic clip + masks, random-init weights",
              "l2": "per-step working set (~4 GB of activations) >> 126 MB L2, no explicit flush",
              "parallelism": f"clips sharded over {world} GPU(s), weights replicated"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        fps, threads, sample, ms = cpu_reference_fps(args.steps, args.warmup)
        line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

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
    frames, masks = D.clips.synthetic_clip(1000 + rank, B, CLIP_FRAMES, H, W)    # one clip per rank
    xin_host = torch.cat([frames, masks], dim=2).contiguous().pin_memory()        # (B,T,4,H,W) pinned
    xin_dev = xin_host.to(dev)
    stats = D.clips.ClipStats(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def qp_at(t):
        return mp.shift_qp(BASE_QP, D.clips.INDEX_MAP[t % 8])

    with torch.no_grad():
        # GOP start (untimed): I-frame, first P-frame (after_i=True)
        r = mi(xin_dev[:, 0, :3].contiguous(), BASE_QP)
        r = mp(xin_dev[:, 1], qp_at(1), r["dpb"], after_i=True)
        dpb = r["dpb"]
        step_no = [1]

        def step_resident():
            nonlocal dpb
            step_no[0] += 1
            t = 1 + (step_no[0] % (CLIP_FRAMES - 1))
            res = mp(xin_dev[:, t], qp_at(step_no[0]), dpb, after_i=False)
            dpb = res["dpb"]
            return res, t

        copy_stream = torch.cuda.Stream(dev)

        def prefetch(step_index):
            """H2D of the frame + mask of step `step_index` from pinned host memory, on the copy stream."""
            t = 1 + (step_index % (CLIP_FRAMES - 1))
            with torch.cuda.stream(copy_stream):
                x = xin_host[:, t].to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return x, ev

        def step_host(x, ev, out_host):
            """One step through the public API on a frame that was copied from the host for this step; its
            (bpp, bpp_y, bpp_z) go back to pinned host memory and an event marks when they are readable."""
            nonlocal dpb
            step_no[0] += 1
            torch.cuda.current_stream().wait_event(ev)
            x.record_stream(torch.cuda.current_stream())
            res = mp(x, qp_at(step_no[0]), dpb, after_i=False)
            dpb = res["dpb"]
            out_host.copy_(torch.stack([res["bpp"], res["bpp_y"], res["bpp_z"]], 1), non_blocking=True)
            done = torch.cuda.Event()
            done.record()
            return done

        def run_host(n):
            """n end-to-end steps, software-pipelined two deep: the H2D copy of step i+1 overlaps the kernels of
            step i, and the host reads the bpp of step i (after its D2H) while step i+1 is already queued."""
            outs = [torch.empty(B, 3).pin_memory() for _ in range(2)]
            total = 0.0
            nxt = prefetch(step_no[0] + 1)
            prev = None
            for i in range(n):
                x, ev = nxt
                done = step_host(x, ev, outs[i & 1])
                if i + 1 < n:
                    nxt = prefetch(step_no[0] + 1)
                if prev is not None:
                    prev[0].synchronize()
                    total += float(prev[1][0, 0])          # the caller consumes bpp of every step
                prev = (done, outs[i & 1])
            prev[0].synchronize()
            total += float(prev[1][0, 0])
            return total

        # W untimed warm-up steps, extended until the device has been busy for ~1 s: a B200 that idled through
        # model construction needs that long to settle its clocks (the first 20 steps of a fresh process measured
        # 6 % slower than the next 20 otherwise)
        import time as _time
        t_warm = _time.time()
        n_warm = 0
        while n_warm < max(3, args.warmup) or _time.time() - t_warm < 1.0:
            step_resident()
            n_warm += 1
            if n_warm % 8 == 0:
                torch.cuda.synchronize(dev)
        torch.cuda.synchronize(dev)
        # ---- timed: resident inputs
        barrier()
        l0 = lib.dmc_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local_rank) as clk:
            if args.profiler_range:
                torch.cuda.profiler.start()
            e0.record()
            for _ in range(args.steps):
                res, t = step_resident()
                stats.add_frame(res, xin_dev[:, t, :3], xin_dev[:, t, 3:4])
            stats.all_reduce()
            e1.record()
            barrier()
            if args.profiler_range:
                torch.cuda.profiler.stop()
        launches = lib.dmc_kernel_launches() - l0
        ms = e0.elapsed_time(e1)
        # ---- timed: end to end through the public API with host buffers
        run_host(2)
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        run_host(args.steps)
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)
        # ---- contraction kernel: device time per launch (CUDA events on the launching stream)
        h, _ = mp._engine(B, H, W, dev)
        lib.dmc_profile_enable(h, 1)
        nprof = 3
        for _ in range(nprof):
            step_resident()
        import ctypes
        g_ms, g_n, g_fl, g_is = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double(), ctypes.c_double()
        lib.dmc_profile_read(h, ctypes.byref(g_ms), ctypes.byref(g_n), ctypes.byref(g_fl), ctypes.byref(g_is))
        lib.dmc_profile_enable(h, 0)

    times = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e = times.tolist()
    frames_done = args.steps * B * world
    value = frames_done / (ms / 1e3)
    e2e = frames_done / (ms_e2e / 1e3)
    if rank == 0:
        pk = peaks()
        per_launch_ms = g_ms.value / max(1, g_n.value)
        achieved = (g_fl.value / max(1, g_n.value)) / (per_launch_ms * 1e-3) / 1e12 if g_n.value else 0.0
        summ = stats.summary()
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 tensor-core operands (tcgen05 kind::f16), fp32 accumulate: 3-term split product (fp16 hi + "
                         "2^11-scaled fp16 lo) on every layer that can reach a symbol; single term in recon_generation_net",
                "data": "synthetic", "config": config, "clocks": clk.summary(),
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": B * 4 * H * W * 4,
                        "d2h_bytes_per_step": B * 3 * 4},
                "gpu_launches": int(launches),
                "roofline": {"bound": "tensor",
                             "kernel": "k_gemm_s3_chain (persistent tcgen05 chain of 1x1 layers, 3-term split-fp16 product)",
                             "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": achieved / pk["bf16_tflops"] if pk["bf16_tflops"] else None,
                             "traffic": (ncu_traffic() or {}).get("bytes"),
                             "traffic_detail": ncu_traffic(),
                             "peak_source": pk["source"],
                             "launches_per_frame": g_n.value / nprof, "avg_launch_ms": per_launch_ms,
                             "gemm_share_of_step": (g_ms.value / nprof) / (ms / args.steps),
                             "issued_mma_tflops": g_is.value / (g_ms.value * 1e-3) / 1e12 if g_ms.value else 0.0,
                             "issued_frac": (g_is.value / (g_ms.value * 1e-3) / 1e12 / pk["bf16_tflops"])
                             if g_ms.value and pk["bf16_tflops"] else None,
                             "note": "achieved = algorithmic conv FLOPs (2*M*N*K) per contraction launch / mean launch "
                                     "time (CUDA events around every launch, 3 frames); a launch is a chain of 1-5 "
                                     "layers.  fp32-grade layers issue 3 fp16 MMA terms per product, so frac <= 1/3 "
                                     "there: issued_mma_tflops / issued_frac count the MMAs actually issued"},
                "quality": {"bpp": summ["bpp"], "psnr": summ["psnr"], "roi_psnr": summ["roi_psnr"],
                            "frames": summ["frames"]},
                "algorithmic_tflops": ALGO_GFLOP_PER_FRAME * value / 1e3}
        if world == 1 and not args.no_cpu_baseline:
            fps, threads, sample, _ = cpu_reference_fps(2, 0, budget_s=25.0)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
