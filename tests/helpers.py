"""Shared test helpers: seeded models, golden access, parity metrics."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dmc_b200 as D                      # noqa: E402  (the product package)
from oracle import dmc_oracle as O        # noqa: E402  (the checker)
from oracle import golden_cases as gc     # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# parity gates of BASELINE.json north_star
SYMBOL_MATCH_MIN = 0.9999
BPP_REL_TOL = 1e-3
PSNR_TOL_DB = 0.02


def golden(name):
    return np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))


def seeded_models(variant, case, device="cpu"):
    """Drop-in modules initialised exactly like the reference under the golden seeds."""
    torch.manual_seed(gc.SEED_I)
    mi = D.DMCI().eval()
    gc.perturb(mi, case)
    torch.manual_seed(gc.SEED_P)
    mp = D.build_p_model(variant).eval()
    gc.perturb(mp, case)
    return mi.to(device), mp.to(device)


def sd_of(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


def symbol_match(a, b):
    a = torch.as_tensor(a).float().flatten()
    b = torch.as_tensor(b).float().flatten()
    return float((a == b).float().mean()), int((a != b).sum())


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float(((a - b).abs() / b.abs().clamp_min(1e-12)).max())
