"""TEST INFRASTRUCTURE ONLY.  Definitions shared by oracle/make_golden.py (which runs the real
reference) and the tests (which replay the same seeded cases through the oracle restatement and
the CUDA engine): case list, input recipe, weight perturbation, what gets recorded."""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VARIANTS = ("old", "performance", "fast", "mask_prop")
INDEX_MAP = (0, 1, 0, 2, 0, 2, 0, 2)
SEED_I, SEED_P = 0, 1

CASES = (
    {"name": "anchor_256", "kind": "rand", "B": 1, "T": 4, "H": 256, "W": 256, "qp": 32, "perturb": False},
    {"name": "rect_128x192", "kind": "clip", "B": 2, "T": 3, "H": 128, "W": 192, "qp": 20, "perturb": True},
    # ragged: y is 5 x 7 and gets replicate-padded to 8 x 8 for the hyper path (models/common_model.py:68-72);
    # `performance` does not pad (seg_video_model.py:331) and cannot run this size in the reference either
    {"name": "ragged_80x112", "kind": "clip", "B": 1, "T": 3, "H": 80, "W": 112, "qp": 28, "perturb": True,
     "variants": ("old", "fast", "mask_prop")},
)


def case_variants(case):
    return case.get("variants", VARIANTS)


def case_by_name(name):
    return next(c for c in CASES if c["name"] == name)


def case_inputs(case):
    """frames (B,T,3,H,W) in [0,1] and masks (B,T,1,H,W) in {0,1}."""
    B, T, H, W = case["B"], case["T"], case["H"], case["W"]
    if case["kind"] == "rand":            # SURVEY.md section 4 recipe
        g = torch.Generator().manual_seed(2)
        frames = torch.rand(B, T, 3, H, W, generator=g)
        masks = torch.zeros(B, T, 1, H, W)
        masks[..., 64:160, 80:200] = 1.0
        return frames, masks
    clips = importlib.import_module("semantic-segmentation-guided-neural-video-compression_b200.clips")
    return clips.synthetic_clip(7, B, T, H, W)


def perturb(model_or_sd, case):
    """Make the per-QP tables non-trivial (they are all-ones at init)."""
    if not case["perturb"]:
        return
    sd = model_or_sd if isinstance(model_or_sd, dict) else dict(model_or_sd.named_parameters())
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for k in sorted(sd):
            if k.startswith("q_"):
                noise = 0.1 * torch.randn(sd[k].shape, generator=g)
                sd[k].add_(noise.to(sd[k].device))


def sd_checksum(sd) -> np.ndarray:
    tot, wtot = 0.0, 0.0
    for i, k in enumerate(sorted(sd)):
        v = sd[k].detach().double().cpu().flatten()
        tot += float(v.abs().sum())
        wtot += float((v * torch.arange(1, v.numel() + 1, dtype=torch.float64)).sum()) * (i + 1)
    return np.array([tot, wtot], dtype=np.float64)


def psnr(mse: float) -> float:
    return float(10.0 * np.log10(1.0 / (mse + 1e-12)))


def metrics(x_hat, target, mask):
    d2 = (x_hat.double() - target.double()) ** 2
    mse = float(d2.mean())
    if mask is not None and float(mask.sum()) > 0:
        m = (mask > 0).double().expand_as(d2)
        roi = float((d2 * m).sum() / m.sum())
    else:
        roi = mse
    return psnr(mse), psnr(roi)


def record(rec, tag, result, box, target, mask):
    """Store the observable outputs of one forward under `tag/...`."""
    x_hat = result["dpb"]["frame"]
    rec[f"{tag}/bpp3"] = torch.stack([result["bpp"], result["bpp_y"], result["bpp_z"]], 1).numpy().astype(np.float32)
    p, r = metrics(x_hat, target, mask)
    rec[f"{tag}/psnr"] = np.array([p, r], dtype=np.float64)
    rec[f"{tag}/x_hat_sub"] = x_hat[:, :, ::8, ::8].numpy().astype(np.float32)
    rec[f"{tag}/x_hat_sum"] = np.array([float(x_hat.double().sum()), float((x_hat.double() ** 2).sum())])
    feat = result["dpb"].get("feature")
    if feat is not None:
        rec[f"{tag}/feature_sub"] = feat[:, :, ::4, ::4].numpy().astype(np.float32)
        rec[f"{tag}/feature_sum"] = np.array([float(feat.double().sum()), float((feat.double() ** 2).sum())])
    y_q = box["y_q"]
    assert float(y_q.abs().max()) < 127 and float(box["z_hat"].abs().max()) < 127
    rec[f"{tag}/y_q"] = y_q.numpy().astype(np.int8)
    rec[f"{tag}/z_hat"] = box["z_hat"].numpy().astype(np.int8)
    rec[f"{tag}/scales_sub"] = box["scales_hat"][:, :, ::2, ::2].numpy().astype(np.float32)
    mp = result.get("mask_pred")
    if mp is not None:
        rec[f"{tag}/mask_pred_sub"] = mp[:, :, ::8, ::8].numpy().astype(np.float32)
