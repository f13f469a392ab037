"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the REAL reference
(imported read-only from /root/reference) on seeded inputs.  Run in the build container:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The fixtures pin oracle/dmc_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA
path.  Weights are not stored: they are the reference's default initialisation under a fixed
torch seed, which the drop-in modules reproduce bit-for-bit; per-model checksums are stored so
that a drift of torch's RNG or init code is detected instead of silently changing the test.

Cases (per P variant):
  anchor  SURVEY.md section 4 known-answer recipe: 256x256, B=1, 1 I + 3 P, torch.rand frames,
          rectangular mask, qp 32  (BASELINE.json config 1 for `old`)
  rect    128x192, B=2, synthetic drifting clip, per-QP tables perturbed so that qp indexing
          matters, 1 I + 2 P, qp 20
  ragged  80x112, B=1: y (5 x 7) is replicate-padded to 8 x 8 for the hyper path; old / fast / mask_prop only

`python oracle/make_golden.py NAME ...` regenerates only the named cases.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.dont_write_bytecode = True
REF = os.environ.get("DMC_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from src.models.image_model import DMCI as RefDMCI                      # noqa: E402
from src.models.video_model import DMC as RefOld                        # noqa: E402
from src.refactor.config import DMCConfig as RefCfg                     # noqa: E402
from src.refactor.mask_prop_seg_video_model import DMC as RefMaskProp   # noqa: E402
from src.refactor.seg_video_model import DMC as RefPerf                 # noqa: E402
from src.refactor.seg_video_model_fast import DMC as RefFast            # noqa: E402

from oracle import golden_cases as gc                                   # noqa: E402

REF_P = {"old": lambda: RefOld(), "performance": lambda: RefPerf(RefCfg()),
         "fast": lambda: RefFast(RefCfg()), "mask_prop": lambda: RefMaskProp(RefCfg())}


def capture(model):
    """Record the arguments of the two likelihood calls of forward()."""
    box = {}
    gy, gz = model.get_y_gaussian_bits, model.get_z_bits

    def wy(y, sigma):
        box["y_q"], box["scales_hat"] = y.clone(), sigma.clone()
        return gy(y, sigma)

    def wz(z, est, idx):
        box["z_hat"] = z.clone()
        return gz(z, est, idx)

    model.get_y_gaussian_bits, model.get_z_bits = wy, wz
    return box


@torch.no_grad()
def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(8)
    only = set(sys.argv[1:])
    for case in gc.CASES:
        if only and case["name"] not in only:
            continue
        frames, masks = gc.case_inputs(case)
        torch.manual_seed(gc.SEED_I)
        ref_i = RefDMCI().eval()
        gc.perturb(ref_i, case)
        rec = {"sd_checksum_intra": gc.sd_checksum(ref_i.state_dict())}
        box_i = capture(ref_i)
        r_i = ref_i(frames[:, 0], case["qp"])
        gc.record(rec, "intra/0", r_i, box_i, frames[:, 0], None)
        for variant in gc.case_variants(case):
            torch.manual_seed(gc.SEED_P)
            ref_p = REF_P[variant]().eval()
            gc.perturb(ref_p, case)
            rec[f"sd_checksum_{variant}"] = gc.sd_checksum(ref_p.state_dict())
            box = capture(ref_p)
            dpb = r_i["dpb"]
            for t in range(1, frames.shape[1]):
                qp = ref_p.shift_qp(case["qp"], gc.INDEX_MAP[t % 8])
                x_in = frames[:, t] if variant == "old" else torch.cat([frames[:, t], masks[:, t]], 1)
                r = ref_p(x_in, qp, dpb, after_i=(t == 1))
                dpb = r["dpb"]
                gc.record(rec, f"{variant}/{t}", r, box, frames[:, t], masks[:, t])
                print(case["name"], variant, t, qp, [round(float(v), 6) for v in r["bpp"]])
        path = os.path.join(out_dir, f"{case['name']}.npz")
        np.savez_compressed(path, **rec)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
