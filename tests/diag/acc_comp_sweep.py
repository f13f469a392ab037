"""Symbol flips of the CUDA path against the oracle at 1920x1280 as a function of the accumulate-truncation
compensation kappa (csrc/kernels.cu: acc_comp_scaled); not a test.

    python tests/diag/acc_comp_sweep.py [variant] [kappa,kappa,...] [frames]

The oracle runs once (intra + P frames, free running on its own dpb).  For every kappa the CUDA path runs
  (a) on identical inputs per call (the oracle's dpb)  and  (b) free running on its own dpb,
and the y symbols that differ from the oracle's are counted per frame (gate: 122 of 1 228 800; intra: 245 of 2 457 600).
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

variant = sys.argv[1] if len(sys.argv) > 1 else "performance"
kappas = [float(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,0.276").split(",")]
T_ = int(sys.argv[3]) if len(sys.argv) > 3 else 4
H, W = 1280, 1920
lib = D._capi.load()

frames, masks = D.clips.synthetic_clip(3, 1, T_, H, W)
torch.manual_seed(gc.SEED_I)
mi0 = D.DMCI().eval()
torch.manual_seed(gc.SEED_P)
mp0 = D.build_p_model(variant).eval()
sd_i, sd_p = sd_of(mi0), sd_of(mp0)


def x_of(t, dev=None):
    x = frames[:, t] if variant == "old" else torch.cat([frames[:, t], masks[:, t]], 1)
    return x.cuda() if dev else x


t0 = time.time()
ora = []
with torch.no_grad():
    ti = {}
    o = O.dmci_forward(sd_i, frames[:, 0], 32, ti)
    ora.append((o, ti["y_q"]))
    dpb = o["dpb"]
    for t in range(1, T_):
        qp = O.shift_qp(32, O.INDEX_MAP[t % 8])
        to = {}
        o = O.dmc_forward(sd_p, variant, x_of(t), qp, dpb, after_i=(t == 1), taps=to)
        ora.append((o, to["y_q"]))
        dpb = o["dpb"]
print(f"oracle: intra + {T_ - 1} P frames in {time.time() - t0:.1f} s", flush=True)


def cuda_dpb(d):
    return {k: (v.cuda() if v is not None else None) for k, v in d.items()}


for kappa in kappas:
    lib.dmc_set_acc_comp(kappa)
    torch.manual_seed(gc.SEED_I)
    mi = D.DMCI().eval().cuda()
    torch.manual_seed(gc.SEED_P)
    mp = D.build_p_model(variant).eval().cuda()
    mi.engine_flags = mp.engine_flags = T.capi.FLAG_KEEP_TAPS
    with torch.no_grad():
        c = mi(frames[:, 0].cuda(), 32)
        _, bad = symbol_match(mi.get_tap("y_q", frames[:, 0].cuda()).cpu(), ora[0][1])
        line = [f"kappa {kappa:5.3f} {variant}: intra {bad:4d} (bpp rel {T.rel_err(c['bpp'].cpu(), ora[0][0]['bpp']):.1e})"]
        dpb_free = c["dpb"]
        for t in range(1, T_):
            qp = mp.shift_qp(32, O.INDEX_MAP[t % 8])
            x = x_of(t, True)
            ca = mp(x, qp, cuda_dpb(ora[t - 1][0]["dpb"]), after_i=(t == 1))
            _, bad_a = symbol_match(mp.get_tap("y_q", x).cpu(), ora[t][1])
            cf = mp(x, qp, dpb_free, after_i=(t == 1))
            _, bad_f = symbol_match(mp.get_tap("y_q", x).cpu(), ora[t][1])
            dpb_free = cf["dpb"]
            line.append(f"P{t} identical-inputs {bad_a:4d} free-running {bad_f:5d} "
                        f"(bpp rel {T.rel_err(cf['bpp'].cpu(), ora[t][0]['bpp']):.1e})")
    print(" | ".join(line), flush=True)
    del mi, mp
    torch.cuda.empty_cache()
