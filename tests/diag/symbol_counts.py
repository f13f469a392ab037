"""Counts quantised-symbol mismatches of the CUDA path against the oracle, frame by frame (not a test).

    python tests/diag/symbol_counts.py [--full] [variant ...]

Golden-case GOPs for every variant; --full adds a free-running 1 I + 3 P GOP at 1920x1280 (the
size the 99.99 % gate of BASELINE.json is quoted on: <= 122 mismatches of 1 228 800 symbols);
--teacher feeds every CUDA forward call the ORACLE's dpb (identical inputs per call) instead of its own.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import test_gpu_parity as T  # noqa: E402
from helpers import D, O, gc, sd_of, symbol_match  # noqa: E402


def golden_cases(variants):
    for variant in variants:
        for case in gc.CASES:
            if variant not in gc.case_variants(case):
                continue
            rep = T._run_case(variant, case, T.capi.FLAG_KEEP_TAPS, record_taps=("y_q", "z_hat"))
            for tag, o, c, to, tc, target, mask in rep:
                line = f"{variant:11s} {case['name']:13s} {tag:5s} bpp rel {T.rel_err(c['bpp'].cpu(), o['bpp']):.2e}"
                if tc:
                    fy, by = symbol_match(tc["y_q"], to["y_q"])
                    line += f"  y bad {by}/{to['y_q'].numel()}"
                if "z_hat" in tc:
                    fz, bz = symbol_match(tc["z_hat"], to["z_hat"])
                    line += f"  z bad {bz}/{to['z_hat'].numel()}"
                if o["dpb"].get("feature") is not None:
                    d = (o["dpb"]["feature"] - c["dpb"]["feature"].cpu()).abs()
                    line += f"  feature max diff {float(d.max()):.2e}"
                print(line, flush=True)


def full_size(variant="performance", H=1280, W=1920, T_=4):
    frames, masks = D.clips.synthetic_clip(3, 1, T_, H, W)
    torch.manual_seed(gc.SEED_I)
    mi = D.DMCI().eval()
    torch.manual_seed(gc.SEED_P)
    mp = D.build_p_model(variant).eval()
    sd_i, sd_p = sd_of(mi), sd_of(mp)
    mi, mp = mi.cuda(), mp.cuda()
    mi.engine_flags = mp.engine_flags = T.capi.FLAG_KEEP_TAPS | (T.capi.FLAG_SIMT_GEMM if "--simt" in sys.argv else 0)
    fr, mk = frames.cuda(), masks.cuda()
    with torch.no_grad():
        t0 = time.time()
        ti = {}
        o_i = O.dmci_forward(sd_i, frames[:, 0], 32, ti)
        c_i = mi(fr[:, 0], 32)
        fy, by = symbol_match(mi.get_tap("y_q", fr[:, 0]).cpu(), ti["y_q"])
        dx = (o_i["dpb"]["frame"] - c_i["dpb"]["frame"].cpu()).abs()
        print(f"full {variant} intra: y bad {by}/{ti['y_q'].numel()}  bpp rel {T.rel_err(c_i['bpp'].cpu(), o_i['bpp']):.2e} "
              f" x_hat max diff {float(dx.max()):.2e} ({time.time() - t0:.1f} s)", flush=True)
        dpb_o, dpb_c = o_i["dpb"], c_i["dpb"]
        if "--teacher" in sys.argv:
            dpb_c = {k: (v.cuda() if v is not None else None) for k, v in o_i["dpb"].items()}
        for t in range(1, T_):
            qp = mp.shift_qp(32, O.INDEX_MAP[t % 8])
            xo = frames[:, t] if variant == "old" else torch.cat([frames[:, t], masks[:, t]], 1)
            xc = fr[:, t] if variant == "old" else torch.cat([fr[:, t], mk[:, t]], 1)
            to = {}
            o = O.dmc_forward(sd_p, variant, xo, qp, dpb_o, after_i=(t == 1), taps=to)
            c = mp(xc, qp, dpb_c, after_i=(t == 1))
            fy, by = symbol_match(mp.get_tap("y_q", xc).cpu(), to["y_q"])
            fz, bz = symbol_match(mp.get_tap("z_hat", xc).cpu(), to["z_hat"])
            po, ro = gc.metrics(o["dpb"]["frame"], frames[:, t], masks[:, t])
            pc, rc = gc.metrics(c["dpb"]["frame"].cpu(), frames[:, t], masks[:, t])
            print(f"full {variant} P{t}: y bad {by}/{to['y_q'].numel()} ({fy:.6f})  z bad {bz}  bpp rel "
                  f"{T.rel_err(c['bpp'].cpu(), o['bpp']):.2e}  dPSNR {abs(po - pc):.1e} dROI {abs(ro - rc):.1e} "
                  f"({time.time() - t0:.1f} s)", flush=True)
            dpb_o, dpb_c = o["dpb"], c["dpb"]
            if "--teacher" in sys.argv:      # identical inputs for every forward call: the CUDA path gets the oracle's dpb
                dpb_c = {k: (v.cuda() if v is not None else None) for k, v in o["dpb"].items()}


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if "--full-only" not in sys.argv:
        golden_cases(args or gc.VARIANTS)
    if "--full" in sys.argv or "--full-only" in sys.argv:
        full_size()
