"""Device side of the caller's data path: what `WaymoCameraYCbCrDataset.__getitem__` does to a
decoded camera frame, on the GPU, plus the mask hand-over of a self-propagating `mask_prop` GOP.

    frames_from_u8     src/dataset/seg_waymo_dataset.py:26-43   uint8 RGB -> [0,1] -> BT.709 YCbCr, clamp
                       src/dataset/seg_waymo_dataset.py:56-79   cached mask -> {0,1}
                       src/dataset/seg_waymo_dataset.py:231-245 one crop for the sequence, mask appended as channel 4
    mask_from_logits   the config-4 protocol of SURVEY.md 8(d): frame t is coded with (mask_pred of frame t-1 > 0)

JPEG decoding stays where the reference has it (cv2.imdecode on the host); what crosses PCIe is then 3 + 1 bytes per
pixel instead of 16, and the conversion runs in one launch (csrc/kernels.cu: k_frames_from_u8), bit-identical to the
reference's CPU fp32 arithmetic.  No CPU path: CUDA tensors required.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _capi

__all__ = ["frames_from_u8", "mask_from_logits"]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


def frames_from_u8(img: torch.Tensor, mask: Optional[torch.Tensor] = None,
                   crop: Optional[Tuple[int, int, int, int]] = None, bgr: bool = False, mask_threshold: int = 0,
                   with_mask: bool = True) -> torch.Tensor:
    """img: (T, H, W, 3) uint8 CUDA, interleaved R,G,B (bgr=True: B,G,R as cv2.imdecode returns it).
    mask: (T, H, W) uint8 CUDA or None (the dataset's `strict_masks=False` fallback: zeros).
    crop: (top, left, height, width) applied to every frame, None = whole frame.
    Returns (T, 4, h, w) float32 [Y, Cb, Cr, mask] (or (T, 3, h, w) with with_mask=False)."""
    if not img.is_cuda:
        raise RuntimeError("dmc_b200.data: CUDA tensors required (this implementation has no CPU path)")
    if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
        raise TypeError(f"img must be uint8 (T, H, W, 3), got {img.dtype} {tuple(img.shape)}")
    T, H, W, _ = img.shape
    img = img.contiguous()
    if mask is not None:
        if mask.dtype != torch.uint8 or tuple(mask.shape) != (T, H, W) or mask.device != img.device:
            # the reference raises on a size mismatch too (seg_waymo_dataset.py:68-69)
            raise ValueError(f"mask must be uint8 {(T, H, W)} on {img.device}, got {mask.dtype} {tuple(mask.shape)}")
        mask = mask.contiguous()
    top, left, h, w = crop if crop is not None else (0, 0, H, W)
    if h > H or w > W or top < 0 or left < 0 or top + h > H or left + w > W:
        raise ValueError(f"crop {crop} exceeds image size {(H, W)}")      # seg_waymo_dataset.py:235-236
    C = 4 if with_mask else 3
    out = torch.empty((T, C, h, w), dtype=torch.float32, device=img.device)
    lib = _capi.load()
    with torch.cuda.device(img.device):
        st = ctypes.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)
        _capi.check(lib.dmc_frames_from_u8(_ptr(img), _ptr(mask), _ptr(out), T, H, W, int(top), int(left), int(h), int(w),
                                           C, 1 if bgr else 0, int(mask_threshold), st))
    return out


def mask_from_logits(logits: torch.Tensor) -> torch.Tensor:
    """(mask_pred > 0).float() in one launch; mask_pred is what `DMC_mask_prop.forward` returns on non-first P frames."""
    if not logits.is_cuda or logits.dtype != torch.float32:
        raise TypeError("dmc_b200.data: float32 CUDA tensor required")
    logits = logits.contiguous()
    out = torch.empty_like(logits)
    if logits.numel() == 0:
        return out
    lib = _capi.load()
    with torch.cuda.device(logits.device):
        st = ctypes.c_void_p(torch.cuda.current_stream(logits.device).cuda_stream)
        _capi.check(lib.dmc_mask_from_logits(_ptr(logits), _ptr(out), logits.numel(), st))
    return out
