"""Caller-side harness around the codec modules: synthetic Waymo-shaped clips, the GOP loop of
the reference's validation_step, per-GPU statistics and clip sharding across ranks.

    GOP loop     trainer_seg_video_model.py:1228-1244 (validation_step)
    qp schedule  trainer:76 (index_map), video_model.py:335-336 (shift_qp)
    metrics      trainer:598-601 (_psnr_from_mse), :655-660 (_roi_mse), :904-934
    sync         trainer:1264-1269 self.log(..., sync_dist=True) -> one all-reduce of a stats vector

Clips are independent (the dpb chain never leaves a clip), so multi-GPU work is sharded by
clip with no data-path collective; only the 7-double statistics vector is all-reduced.
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

INDEX_MAP = (0, 1, 0, 2, 0, 2, 0, 2)
STAT_NAMES = ("bits_y", "bits_z", "sq_err", "roi_sq_err", "roi_elems", "elems", "frames")


def synthetic_clip(seed: int, batch: int, frames: int, height: int, width: int,
                   device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """Deterministic YCbCr clip in [0,1], (B,T,3,H,W), and binary masks (B,T,1,H,W).

    Low-pass filtered uniform noise that drifts a few pixels per frame (so consecutive frames
    correlate like camera motion) and 3-8 rectangles covering ~10-20 % of the frame that drift
    with it -- the value ranges of seg_waymo_dataset.py:36-43 without needing the dataset.
    """
    g = torch.Generator().manual_seed(int(seed))
    pad = 4 * frames + 16
    canvas = torch.rand(batch, 3, height + pad, width + pad, generator=g)
    canvas = F.avg_pool2d(canvas, 9, stride=1, padding=4, count_include_pad=False)
    lo = canvas.amin(dim=(2, 3), keepdim=True)
    hi = canvas.amax(dim=(2, 3), keepdim=True)
    canvas = (canvas - lo) / (hi - lo + 1e-12)
    mcan = torch.zeros(batch, 1, height + pad, width + pad)
    for b in range(batch):
        n_rect = int(torch.randint(3, 9, (1,), generator=g))
        target = float(torch.empty(1).uniform_(0.10, 0.20, generator=g))
        for _ in range(n_rect):
            area = target * height * width / n_rect
            aspect = float(torch.empty(1).uniform_(0.5, 2.0, generator=g))
            rh = max(8, min(height // 2, int(math.sqrt(area / aspect))))
            rw = max(8, min(width // 2, int(area / rh)))
            y0 = int(torch.randint(0, height + pad - rh, (1,), generator=g))
            x0 = int(torch.randint(0, width + pad - rw, (1,), generator=g))
            mcan[b, :, y0:y0 + rh, x0:x0 + rw] = 1.0
    step = torch.randint(-3, 4, (frames, 2), generator=g)
    pos = torch.cumsum(step, 0) + pad // 2
    pos = pos.clamp(0, pad)
    fr, mk = [], []
    for t in range(frames):
        y0, x0 = int(pos[t, 0]), int(pos[t, 1])
        fr.append(canvas[:, :, y0:y0 + height, x0:x0 + width])
        mk.append(mcan[:, :, y0:y0 + height, x0:x0 + width])
    return torch.stack(fr, 1).contiguous().to(device), torch.stack(mk, 1).contiguous().to(device)


def shard_clips(n_clips: int, rank: int, world_size: int) -> List[int]:
    """clip c -> rank c mod world_size."""
    return [c for c in range(n_clips) if c % world_size == rank]


def gop_qp(qp: int, t: int, qp_shift=(0, 8, 4)) -> int:
    return qp + qp_shift[INDEX_MAP[t % 8]]


class ClipStats:
    """Device-resident accumulator of the 7-entry statistics vector (dmc_frame_stats)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.vec = torch.zeros(7, dtype=torch.float64, device=self.device)

    def add_frame(self, result: dict, target: torch.Tensor, mask: Optional[torch.Tensor] = None,
                  bpp3: Optional[torch.Tensor] = None):
        """Adds one forward result.  CUDA tensors go through the fused kernel; CPU tensors (gloo
        tests of the host logic) use the same formulas in torch."""
        x_hat = result["dpb"]["frame"]
        B, _, H, W = x_hat.shape
        if bpp3 is None:
            bpp3 = torch.stack([result["bpp"], result["bpp_y"], result["bpp_z"]], dim=1).contiguous()
        if x_hat.is_cuda:
            from . import _capi
            lib = _capi.load()
            st = ctypes.c_void_p(torch.cuda.current_stream(x_hat.device).cuda_stream)
            target = target.contiguous()
            m = mask.contiguous() if mask is not None else None
            rc = lib.dmc_frame_stats(ctypes.c_void_p(self.vec.data_ptr()), ctypes.c_void_p(x_hat.data_ptr()),
                                     ctypes.c_void_p(target.data_ptr()),
                                     ctypes.c_void_p(m.data_ptr()) if m is not None else ctypes.c_void_p(),
                                     ctypes.c_void_p(bpp3.data_ptr()), B, H, W, st)
            _capi.check(rc, None)
            return
        d2 = ((x_hat - target) ** 2).double()
        v = torch.zeros(7, dtype=torch.float64)
        v[0] = (bpp3[:, 1].double() * H * W).sum()
        v[1] = (bpp3[:, 2].double() * H * W).sum()
        v[2] = d2.sum()
        if mask is not None:
            m = (mask > 0).double().expand_as(d2)
            v[3] = (d2 * m).sum()
            v[4] = m.sum()
        v[5] = d2.numel()
        v[6] = B
        self.vec += v.to(self.vec.device)

    def all_reduce(self):
        """One sum all-reduce of the stats vector (NCCL on GPUs, gloo in CPU tests)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.vec, op=dist.ReduceOp.SUM)
        return self

    def summary(self) -> Dict[str, float]:
        v = self.vec.detach().cpu().tolist()
        pixels = v[5] / 3.0 if v[5] else float("nan")
        mse = v[2] / v[5] if v[5] else float("nan")
        roi_mse = v[3] / v[4] if v[4] > 0 else mse
        psnr = lambda m: 10.0 * math.log10(1.0 / (m + 1e-12))
        return {"frames": v[6], "bpp": (v[0] + v[1]) / pixels, "bpp_y": v[0] / pixels, "bpp_z": v[1] / pixels,
                "mse": mse, "psnr": psnr(mse), "roi_mse": roi_mse, "roi_psnr": psnr(roi_mse)}


@torch.no_grad()
def run_gop(i_model, p_model, variant: str, frames: torch.Tensor, masks: Optional[torch.Tensor], qp: int,
            stats: Optional[ClipStats] = None, mask_feedback: bool = False, i_result: Optional[dict] = None):
    """frames (B,T,3,H,W); frame 0 through the intra model, frames 1.. through the P model.
    `mask_feedback` is the config-4 protocol for mask_prop: frame 2 re-uses mask 1,
    frames >= 3 get the thresholded prediction of the previous frame."""
    res = i_result if i_result is not None else i_model(frames[:, 0], qp)
    dpb = res["dpb"]
    outs = [res]
    prev_pred = None
    for t in range(1, frames.shape[1]):
        cq = p_model.shift_qp(qp, INDEX_MAP[t % 8])
        m = None
        if variant != "old" and masks is not None:
            m = masks[:, t]
            if mask_feedback and variant == "mask_prop":
                if t == 2:
                    m = masks[:, 1]
                elif t >= 3 and prev_pred is not None:
                    m = (prev_pred > 0).float()
        x_in = frames[:, t] if m is None else torch.cat([frames[:, t], m], dim=1)
        res = p_model(x_in, cq, dpb, after_i=(t == 1))
        prev_pred = res.get("mask_pred")
        dpb = res["dpb"]
        if stats is not None:
            stats.add_frame(res, frames[:, t], masks[:, t] if masks is not None else None)
        outs.append(res)
    return outs


class GopCoder:
    """A `mask_prop` GOP that needs ONE segmentation mask (the replacement of per-frame mask
    generation, src/utils/build_cache.py:143-236, by propagation).  Protocol of SURVEY.md 8(d) config 4: P frame 1
    (after the intra frame) and P frame 2 are coded with the given mask; from frame 3 on the mask is the previous
    frame's `mask_pred` (src/refactor/mask_predictor.py:27-46) thresholded at logit 0 (`data.mask_from_logits`).
    Same results as `run_gop(..., mask_feedback=True)`; `masks_used` keeps what each P frame saw."""

    def __init__(self, i_model, p_model, qp: int):
        if getattr(p_model, "variant", None) != "mask_prop":
            raise ValueError("GopCoder needs the mask_prop variant (the only one with a MaskPredictor)")
        self.i_model, self.p_model, self.qp = i_model, p_model, int(qp)
        self.masks_used: List[torch.Tensor] = []

    @torch.no_grad()
    def code(self, frames: torch.Tensor, first_mask: torch.Tensor, stats: Optional[ClipStats] = None):
        """frames (B,T,3,H,W); first_mask (B,1,H,W) in {0,1}: the segmentation of frame 1."""
        from . import data
        self.masks_used = []
        res = self.i_model(frames[:, 0], self.qp)
        outs, dpb, pred = [res], res["dpb"], None
        for t in range(1, frames.shape[1]):
            m = first_mask if t <= 2 or pred is None else data.mask_from_logits(pred)
            self.masks_used.append(m)
            res = self.p_model(torch.cat([frames[:, t], m], dim=1), self.p_model.shift_qp(self.qp, INDEX_MAP[t % 8]), dpb,
                               after_i=(t == 1))
            pred, dpb = res.get("mask_pred"), res["dpb"]
            if stats is not None:
                stats.add_frame(res, frames[:, t], m)
            outs.append(res)
        return outs
