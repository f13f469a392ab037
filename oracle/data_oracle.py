"""TEST INFRASTRUCTURE -- CPU restatement of the reference's per-frame data preparation (not part of the product;
only tests/ import it).  Pinned: tests/golden/data_path.npz was produced by the reference's OWN two functions
(`oracle/make_golden_data.py` executes their source from /root/reference; the module itself cannot be imported
here because cv2 and the Waymo reader are absent), and tests/test_oracle_golden.py replays it through this file.

    rgb_from_u8          src/dataset/seg_waymo_dataset.py:26-34   torch.as_tensor(rgb, float32).permute(2,0,1) / 255.0
    rgb_to_ycbcr_bt709   src/dataset/seg_waymo_dataset.py:36-43
    mask_to_float        src/dataset/seg_waymo_dataset.py:56-79   npz cache: 0/1 as stored; png cache: > 127
    item                 src/dataset/seg_waymo_dataset.py:231-245 crop, mask as channel 4, stack over the sequence
"""
from __future__ import annotations

import numpy as np
import torch


def rgb_from_u8(rgb_hwc: np.ndarray) -> torch.Tensor:
    return torch.as_tensor(rgb_hwc, dtype=torch.float32).permute(2, 0, 1) / 255.0


def rgb_to_ycbcr_bt709(rgb_chw: torch.Tensor) -> torch.Tensor:
    r, g, b = rgb_chw[0], rgb_chw[1], rgb_chw[2]
    Kr, Kg, Kb = 0.2126, 0.7152, 0.0722
    y = Kr * r + Kg * g + Kb * b
    cb = 0.5 * (b - y) / (1 - Kb) + 0.5
    cr = 0.5 * (r - y) / (1 - Kr) + 0.5
    return torch.stack([y, cb, cr], dim=0).clamp(0.0, 1.0)


def mask_to_float(mask_hw: np.ndarray, threshold: int = 0) -> torch.Tensor:
    return torch.from_numpy((mask_hw > threshold).astype(np.float32))[None, ...]


def item(img_thwc: np.ndarray, mask_thw, crop=None, bgr=False, threshold=0) -> torch.Tensor:
    """(T, H, W, 3) uint8 [+ (T, H, W) uint8] -> (T, 4, h, w) float32 like `__getitem__`'s second output."""
    out = []
    for t in range(img_thwc.shape[0]):
        rgb = img_thwc[t][..., ::-1].copy() if bgr else img_thwc[t]
        y = rgb_to_ycbcr_bt709(rgb_from_u8(rgb))
        m = mask_to_float(mask_thw[t], threshold) if mask_thw is not None else torch.zeros(1, *y.shape[1:])
        if crop is not None:
            top, left, h, w = crop
            y, m = y[:, top:top + h, left:left + w], m[:, top:top + h, left:left + w]
        out.append(torch.cat([y, m], dim=0))
    return torch.stack(out, dim=0)
