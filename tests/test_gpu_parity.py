"""Parity of the CUDA path (through the C ABI) with the oracle, on a B200.

Gates (BASELINE.json north_star): >= 99.99 % of quantised y symbols identical, bpp within 1e-3
relative, PSNR / ROI-PSNR within 0.02 dB.  Layer-level checks compare single operators with
torch fp32 on the CPU (TF32 never involved); tolerances are written next to each check.
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import (BPP_REL_TOL, D, O, PSNR_TOL_DB, SYMBOL_MATCH_MIN, gc, golden, rel_err, sd_of, seeded_models,
                     symbol_match)

pytestmark = pytest.mark.gpu
capi = D._capi
BACKENDS = {"umma": 0, "simt": 1}


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


def op_conv2d(x, w, b, stride=1, padding=0, groups=1, act=0, nsplit=3, backend=0):
    lib = capi.load()
    xb, wb = x.cuda().contiguous(), w.cuda().contiguous()
    bb = b.cuda().contiguous() if b is not None else None
    B, cin, H, W = x.shape
    cout, _, k, _ = w.shape
    Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
    out = torch.empty(B, cout, Ho, Wo, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.dmc_op_conv2d(_p(xb), _p(wb), _p(bb), _p(out), B, cin, H, W, cout, k, stride, padding, groups, act,
                           nsplit, backend, st)
    capi.check(rc, None)
    return out.cpu()


def _act(y, act):
    return O.wsilu(y) if act == 1 else (F.relu(y) if act == 2 else y)


CONV_CASES = [
    # cin, cout, k, stride, pad, H, W, act
    (256, 256, 1, 1, 0, 32, 48, 0),      # plain 1x1 (dc.0 / dc.3 shape)
    (256, 256, 1, 1, 0, 32, 48, 1),      # + WSiLU epilogue
    (192, 256, 1, 1, 0, 16, 24, 0),      # encoder.conv1
    (512, 256, 1, 1, 0, 16, 24, 0),      # adaptor over a concat
    (320, 192, 1, 1, 0, 16, 24, 0),      # recon head (N = 1.5 tiles)
    (384, 384, 1, 1, 0, 8, 12, 0),       # prior fusion
    (368, 368, 1, 1, 0, 8, 12, 0),       # DMCI width: K and N not multiples of 64
    (256, 128, 3, 2, 1, 32, 48, 0),      # encoder.down (3x3 s2 p1)
    (128, 128, 2, 2, 0, 16, 24, 0),      # hyper down (2x2 s2)
    (64, 64, 3, 1, 1, 16, 24, 1),        # mask predictor 3x3 p1 + WSiLU
    (512, 514, 1, 1, 0, 8, 12, 0),       # DMCI y_prior_fusion.3 (odd N -> SIMT route)
    (514, 256, 1, 1, 0, 8, 12, 0),       # DMCI reduction (odd K -> SIMT route)
]


@pytest.mark.parametrize("backend", list(BACKENDS))
@pytest.mark.parametrize("cfg", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv2d_matches_torch_fp32(cfg, backend):
    cin, cout, k, s, p, H, W, act = cfg
    g = torch.Generator().manual_seed(cin * 7 + cout + k)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = _act(F.conv2d(x, w, b, stride=s, padding=p), act)
    out = op_conv2d(x, w, b, s, p, 1, act, 3, BACKENDS[backend])
    # fp32-grade contraction (3-term split-fp16 product or fp32 FMA): error ~ sqrt(K) * 2^-23 of the scale
    tol = 2e-5 * float(ref.abs().max())
    assert float((out - ref).abs().max()) <= tol


def test_conv2d_single_term_is_fp16_accurate():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 256, 16, 24, generator=g)
    w = torch.randn(256, 256, 1, 1, generator=g) / 16
    ref = F.conv2d(x, w, None)
    out = op_conv2d(x, w, None, nsplit=1, backend=0)
    err = float((out - ref).abs().max())
    assert 1e-5 < err <= 4e-3 * float(ref.abs().max())      # fp16 hi planes only: ~2^-11 relative, and NOT fp32-exact


def test_depthwise_matches_torch_fp32():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 320, 16, 24, generator=g)
    w = torch.randn(320, 1, 3, 3, generator=g) / 3
    b = torch.randn(320, generator=g)
    ref = F.conv2d(x, w, b, padding=1, groups=320)
    out = op_conv2d(x, w, b, 1, 1, 320)
    assert float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("backend", list(BACKENDS))
@pytest.mark.parametrize("cin,cout,shortcut,use_q", [(256, 256, False, False), (512, 256, False, True),
                                                      (128, 128, True, False), (256, 320, False, True),
                                                      (368, 368, False, False)])
def test_depth_conv_block_matches_oracle(cin, cout, shortcut, use_q, backend):
    _check_dcb(cin, cout, shortcut, use_q, BACKENDS[backend], 2, 16, 24)


def _check_dcb(cin, cout, shortcut, use_q, backend, batch, h, w):
    g = torch.Generator().manual_seed(cin + cout)
    m = D.modules._dcb(cin, cout)
    sd = {"b." + k: v.detach() for k, v in m.state_dict().items()}
    x = torch.randn(batch, cin, h, w, generator=g)
    q = (1 + 0.1 * torch.randn(cout, generator=g)) if use_q else None
    ref = O.depth_conv_block(sd, "b", x, shortcut=shortcut, quant_step=q.view(1, -1, 1, 1) if use_q else None)
    names = ["adaptor", "dc.0", "dc.2", "dc.3", "ffn.0", "ffn.2"]
    keep, ptrs = [], (ctypes.c_void_p * 12)()
    for i, n in enumerate(names):
        for j, part in enumerate(("weight", "bias")):
            t = sd.get(f"b.{n}.{part}")
            if t is not None:
                t = t.cuda().contiguous()
                keep.append(t)
                ptrs[2 * i + j] = t.data_ptr()
    xb = x.cuda()
    qb = q.cuda() if use_q else None
    out = torch.empty(batch, cout, h, w, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = capi.load().dmc_op_depth_conv_block(_p(xb), ptrs, _p(qb), _p(out), batch, cin, cout, h, w, int(shortcut), 3,
                                             backend, st)
    capi.check(rc, None)
    assert float((out.cpu() - ref).abs().max()) <= 3e-5 * float(ref.abs().max())


@pytest.mark.parametrize("cin,cout,h,w", [(256, 256, 160, 240), (192, 320, 160, 240), (384, 384, 80, 120)])
def test_depth_conv_block_chain_at_frame_scale(cin, cout, h, w):
    """The chained launch (dc.3 -> ffn.0 -> ffn.2 in one persistent kernel with row-tile dependencies) on the
    feature-map sizes of a 1920x1280 frame: 150 / 38 row tiles over 74 CTA pairs, random data, so a tile that
    started before its producers were stored would show up as a wrong value."""
    _check_dcb(cin, cout, False, True, 0, 1, h, w)


@pytest.mark.parametrize("formula", [0, 1])
def test_gaussian_bits_match_oracle(formula):
    g = torch.Generator().manual_seed(3)
    n = 1 << 16
    sym = torch.round(torch.randn(n, generator=g) * 3)
    sigma = torch.randn(n, generator=g) * 2          # about half negative, like random-init nets
    sigma[:8] = torch.tensor([float("nan"), float("inf"), -float("inf"), 0.0, 1e-7, 1e12, -3.0, 0.11])
    sym[8:12] = torch.tensor([20.0, -20.0, 0.0, 6.0])
    ref = (O.gaussian_bits_refactor(sym.clamp(-6, 6), sigma) if formula else
           O.gaussian_bits_old(sym, sigma.nan_to_num(1e-5)))
    if not formula:
        sigma = sigma.nan_to_num(1e-5)              # formula 0 has no NaN guard in the reference either
    out = torch.empty(n, device="cuda")
    sb, gb = sym.cuda(), sigma.cuda()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(capi.load().dmc_op_gaussian_bits(_p(sb), _p(gb), _p(out), n, formula, st), None)
    out = out.cpu()
    # The kernel evaluates erf as the reference's CPU path does (correctly rounded from fp64, saturated to +-1 from
    # |x| >= 3.832507 like MKL's vsErf).  What remains: vsErf differs from the correctly rounded value by one ulp on
    # ~5 % of its inputs, and for tail symbols p is a difference of two erf values next to +-1, so one ulp is a whole
    # quantum of p (3e-8).  A CPU emulation of the kernel's formulas against the oracle on these inputs gives:
    # 0.09-0.10 % of the elements off by > 1e-3 bit, max 0.0086 bit (formula 0) / 1.0 bit (formula 1: a quantum next
    # to the 1e-9 floor), sums within 7e-8 / 1.9e-6 relative.  Gates = those figures with a 3x margin.
    d = (out - ref).abs()
    assert float((d > 1e-3).float().mean()) <= 3e-3
    tot, tot_ref = float(out.double().sum()), float(ref.double().sum())
    if formula == 0:
        assert float(d.max()) <= 2.5e-2
        assert abs(tot - tot_ref) <= 1e-6 * tot_ref
    else:
        assert float(d.max()) <= 2.0 and float((d > 0.5).float().mean()) <= 1e-3
        assert abs(tot - tot_ref) <= 1e-5 * tot_ref


# ----------------------------------------------------------------------------------------------
# whole frames
# ----------------------------------------------------------------------------------------------
def _to_cuda(dpb):
    return {k: (v.cuda() if v is not None else None) for k, v in dpb.items()}


def _run_case(variant, case, flags, record_taps=(), resync=True):
    """A GOP through the oracle and through the CUDA path.  The CUDA path runs on its OWN dpb (free-running,
    like the reference's validation loop) as long as its symbols are identical to the oracle's; a frame with
    a flipped symbol (allowed by the 99.99 % gate) perturbs the decoded feature around it, so after such a
    frame the next call gets the oracle's dpb again: every call is judged on identical inputs."""
    frames, masks = gc.case_inputs(case)
    mi, mp = seeded_models(variant, case)
    sd_i, sd_p = sd_of(mi), sd_of(mp)
    mi, mp = mi.cuda(), mp.cuda()
    mi.engine_flags = mp.engine_flags = flags
    fr, mk = frames.cuda(), masks.cuda()
    report = []
    with torch.no_grad():
        ti = {}
        o_i = O.dmci_forward(sd_i, frames[:, 0], case["qp"], ti)
        c_i = mi(fr[:, 0], case["qp"])
        taps_ci = {"y_q": mi.get_tap("y_q", fr[:, 0]).cpu()} if "y_q" in record_taps else {}
        report.append(("intra", o_i, c_i, ti if taps_ci else {}, taps_ci, frames[:, 0], None))
        dpb_o, dpb_c = o_i["dpb"], c_i["dpb"]
        if resync and taps_ci and not torch.equal(taps_ci["y_q"], ti["y_q"]):
            dpb_c = _to_cuda(dpb_o)
        for t in range(1, frames.shape[1]):
            qp = mp.shift_qp(case["qp"], O.INDEX_MAP[t % 8])
            if variant == "old":
                xo, xc = frames[:, t], fr[:, t]
            else:
                xo, xc = torch.cat([frames[:, t], masks[:, t]], 1), torch.cat([fr[:, t], mk[:, t]], 1)
            taps_o = {}
            o = O.dmc_forward(sd_p, variant, xo, qp, dpb_o, after_i=(t == 1), taps=taps_o)
            c = mp(xc, qp, dpb_c, after_i=(t == 1))
            taps_c = {n: mp.get_tap(n, xc).cpu() for n in record_taps}
            report.append((f"P{t}", o, c, taps_o, taps_c, frames[:, t], masks[:, t]))
            dpb_o, dpb_c = o["dpb"], c["dpb"]
            if resync and "y_q" in taps_c and not torch.equal(taps_c["y_q"], taps_o["y_q"]):
                dpb_c = _to_cuda(dpb_o)
    return report


def _bpp_tol(case):
    return BPP_REL_TOL


def _assert_frame(tag, o, c, target, mask, exact_symbols=True, bpp_tol=BPP_REL_TOL):
    for k in ("bpp", "bpp_y", "bpp_z"):
        assert rel_err(c[k].cpu(), o[k]) <= bpp_tol, (tag, k, c[k].cpu(), o[k])
    xo, xc = o["dpb"]["frame"], c["dpb"]["frame"].cpu()
    assert float(xc.min()) >= 0.0 and float(xc.max()) <= 1.0
    po, ro = gc.metrics(xo, target, mask)
    pc, rc = gc.metrics(xc, target, mask)
    assert abs(po - pc) <= PSNR_TOL_DB and abs(ro - rc) <= PSNR_TOL_DB, (tag, po, pc, ro, rc)
    if o["dpb"].get("feature") is not None and exact_symbols:
        # (a flipped symbol moves the decoded feature around it by ~1e-2: only frames without one are compared)
        fo, fc = o["dpb"]["feature"], c["dpb"]["feature"].cpu()
        assert float((fo - fc).abs().max()) <= 1e-3 * max(1.0, float(fo.abs().max())), tag


@pytest.mark.parametrize("backend", list(BACKENDS))
@pytest.mark.parametrize("case", gc.CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("variant", gc.VARIANTS)
def test_gop_parity_with_oracle(variant, case, backend):
    if variant not in gc.case_variants(case):
        pytest.skip("the reference does not pad y in this variant: it cannot run this size")
    flags = capi.FLAG_KEEP_TAPS | (capi.FLAG_SIMT_GEMM if backend == "simt" else 0)
    rep = _run_case(variant, case, flags, record_taps=("y_q", "z_hat", "scales_hat"))
    for tag, o, c, taps_o, taps_c, target, mask in rep:
        exact = not taps_c or torch.equal(taps_c["y_q"], taps_o["y_q"])
        _assert_frame(f"{variant}/{case['name']}/{backend}/{tag}", o, c, target, mask, exact_symbols=exact,
                      bpp_tol=_bpp_tol(case))
        if taps_c:
            frac, bad = symbol_match(taps_c["y_q"], taps_o["y_q"])
            assert frac >= SYMBOL_MATCH_MIN, (tag, "y symbols", frac, bad)
        if "z_hat" in taps_c:
            frac_z, bad_z = symbol_match(taps_c["z_hat"], taps_o["z_hat"])
            assert frac_z >= SYMBOL_MATCH_MIN, (tag, "z symbols", frac_z, bad_z)
        if "scales_hat" in taps_c and exact:
            # the sigma handed to the likelihood (merged over the two checkerboard steps, raw network output)
            so, sc = taps_o["scales_hat"], taps_c["scales_hat"]
            assert float((so - sc).abs().max()) <= 1e-4 * max(1.0, float(so.abs().max())), (tag, "scales_hat")
        if "mask_pred" in o and o["mask_pred"] is not None:
            mo, mc = o["mask_pred"], c["mask_pred"].cpu()
            assert float((mo - mc).abs().max()) <= 1e-4 * max(1.0, float(mo.abs().max()))


@pytest.mark.parametrize("variant", gc.VARIANTS)
def test_cuda_path_against_reference_golden(variant):
    """Straight against the fixtures minted from the real reference (no oracle in between)."""
    case = gc.case_by_name("anchor_256")
    g = golden(case["name"])
    frames, masks = gc.case_inputs(case)
    mi, mp = seeded_models(variant, case)
    mi, mp = mi.cuda(), mp.cuda()
    mp.engine_flags = capi.FLAG_KEEP_TAPS
    fr, mk = frames.cuda(), masks.cuda()
    with torch.no_grad():
        r = mi(fr[:, 0], case["qp"])
        assert rel_err(r["bpp"].cpu(), g["intra/0/bpp3"][:, 0]) <= BPP_REL_TOL
        dpb = r["dpb"]
        for t in range(1, frames.shape[1]):
            qp = mp.shift_qp(case["qp"], O.INDEX_MAP[t % 8])
            x_in = fr[:, t] if variant == "old" else torch.cat([fr[:, t], mk[:, t]], 1)
            r = mp(x_in, qp, dpb, after_i=(t == 1))
            dpb = r["dpb"]
            tag = f"{variant}/{t}"
            assert rel_err(r["bpp"].cpu(), g[f"{tag}/bpp3"][:, 0]) <= BPP_REL_TOL
            frac, bad = symbol_match(mp.get_tap("y_q", x_in).cpu().numpy().astype(np.int8), g[f"{tag}/y_q"])
            assert frac >= SYMBOL_MATCH_MIN, (tag, frac, bad)
            p, roi = gc.metrics(r["dpb"]["frame"].cpu(), frames[:, t], masks[:, t])
            assert abs(p - g[f"{tag}/psnr"][0]) <= PSNR_TOL_DB and abs(roi - g[f"{tag}/psnr"][1]) <= PSNR_TOL_DB


def test_clip_stats_kernel_matches_cpu_formulas():
    g = torch.Generator().manual_seed(9)
    x = torch.rand(2, 3, 128, 192, generator=g)
    xh = (x + 0.02 * torch.randn(2, 3, 128, 192, generator=g)).clamp(0, 1)
    mask = (torch.rand(2, 1, 128, 192, generator=g) > 0.7).float()
    res = {"dpb": {"frame": xh}, "bpp": torch.tensor([1.5, 2.5]), "bpp_y": torch.tensor([1.0, 2.0]),
           "bpp_z": torch.tensor([0.5, 0.5])}
    cpu = D.clips.ClipStats("cpu")
    cpu.add_frame(res, x, mask)
    gpu = D.clips.ClipStats("cuda")
    res_g = {"dpb": {"frame": xh.cuda()}, "bpp": res["bpp"].cuda(), "bpp_y": res["bpp_y"].cuda(),
             "bpp_z": res["bpp_z"].cuda()}
    gpu.add_frame(res_g, x.cuda(), mask.cuda())
    assert torch.allclose(gpu.vec.cpu(), cpu.vec, rtol=1e-6, atol=1e-9)


def test_full_size_properties():
    """1920x1280 (BASELINE.json config 2): properties that need no CPU oracle run."""
    B, H, W = 1, 1280, 1920
    frames, masks = D.clips.synthetic_clip(3, B, 3, H, W)
    torch.manual_seed(gc.SEED_P)
    mp = D.build_p_model("performance").eval().cuda()
    fr, mk = frames.cuda(), masks.cuda()
    with torch.no_grad():
        dpb = {"frame": fr[:, 0], "feature": None}
        x1 = torch.cat([fr[:, 1], mk[:, 1]], 1)
        r1 = mp(x1, 40, dpb, after_i=True)
        r1b = mp(x1, 40, dpb, after_i=True)
        r2 = mp(torch.cat([fr[:, 2], mk[:, 2]], 1), 32, r1["dpb"], after_i=False)
    for r in (r1, r2):
        xh = r["dpb"]["frame"]
        assert xh.shape == (B, 3, H, W) and r["dpb"]["feature"].shape == (B, 256, H // 8, W // 8)
        assert bool(torch.isfinite(xh).all()) and float(xh.min()) >= 0 and float(xh.max()) <= 1
        assert bool(torch.isfinite(r["dpb"]["feature"]).all())
        assert 0 < float(r["bpp"]) < 32 and abs(float(r["bpp"] - r["bpp_y"] - r["bpp_z"])) < 1e-5
    # same inputs -> same bits (run-to-run deterministic apart from fp64 atomics far below fp32 ulp)
    assert torch.equal(r1["dpb"]["frame"], r1b["dpb"]["frame"])
    assert abs(float(r1["bpp"] - r1b["bpp"])) <= 1e-6 * float(r1["bpp"])


_FULL = {}


def _full_size_intra():
    """The 1920x1280 clip and its intra frame through the oracle and the CUDA path (computed once per session)."""
    if not _FULL:
        H, W = 1280, 1920
        frames, masks = D.clips.synthetic_clip(3, 1, 3, H, W)
        torch.manual_seed(gc.SEED_I)
        mi = D.DMCI().eval()
        sd_i = sd_of(mi)
        mi = mi.cuda()
        mi.engine_flags = capi.FLAG_KEEP_TAPS
        with torch.no_grad():
            ti = {}
            o = O.dmci_forward(sd_i, frames[:, 0], 32, ti)
            c = mi(frames[:, 0].cuda(), 32)
            y_c = mi.get_tap("y_q", frames[:, 0].cuda()).cpu()
        _FULL.update(frames=frames, masks=masks, o=o, c={k: (v.cpu() if torch.is_tensor(v) else v) for k, v in c.items()
                                                         if k != "dpb"},
                     x_hat_c=c["dpb"]["frame"].cpu(), y_o=ti["y_q"], y_c=y_c)
        del mi
        torch.cuda.empty_cache()
    return _FULL


def test_full_size_intra_parity():
    """1920x1280 intra frame (DMCI) against the oracle: 2 457 600 symbols, gate >= 99.99 %."""
    f = _full_size_intra()
    frac, bad = symbol_match(f["y_c"], f["y_o"])
    assert frac >= SYMBOL_MATCH_MIN, ("intra", frac, bad)
    assert rel_err(f["c"]["bpp"], f["o"]["bpp"]) <= BPP_REL_TOL
    po, _ = gc.metrics(f["o"]["dpb"]["frame"], f["frames"][:, 0], None)
    pc, _ = gc.metrics(f["x_hat_c"], f["frames"][:, 0], None)
    assert abs(po - pc) <= PSNR_TOL_DB


@pytest.mark.parametrize("variant", gc.VARIANTS)
def test_full_size_parity_on_identical_inputs(variant):
    """1920x1280 (the size of BASELINE.json configs 2-5), every variant: two P frames (after_i True / False)
    against the oracle, EVERY call on identical inputs (the CUDA path gets the oracle's dpb).  This is the size
    the gates are quoted on: >= 99.99 % symbols = at most 122 of 1 228 800 per P frame.  (Free-running, the
    handful of intra-frame flips -- well inside the gate -- are amplified chaotically by the feature recurrence,
    for the fp32-FMA backend as well: tests/diag/symbol_counts.py)"""
    f = _full_size_intra()
    frames, masks = f["frames"], f["masks"]
    torch.manual_seed(gc.SEED_P)
    mp = D.build_p_model(variant).eval()
    sd_p = sd_of(mp)
    mp = mp.cuda()
    mp.engine_flags = capi.FLAG_KEEP_TAPS
    with torch.no_grad():
        dpb_o = f["o"]["dpb"]
        for t in (1, 2):
            qp = mp.shift_qp(32, O.INDEX_MAP[t % 8])
            xo = frames[:, t] if variant == "old" else torch.cat([frames[:, t], masks[:, t]], 1)
            to = {}
            o = O.dmc_forward(sd_p, variant, xo, qp, dpb_o, after_i=(t == 1), taps=to)
            c = mp(xo.cuda(), qp, _to_cuda(dpb_o), after_i=(t == 1))
            frac, bad = symbol_match(mp.get_tap("y_q", xo.cuda()).cpu(), to["y_q"])
            assert frac >= SYMBOL_MATCH_MIN, (t, "y symbols", frac, bad)
            frac_z, bad_z = symbol_match(mp.get_tap("z_hat", xo.cuda()).cpu(), to["z_hat"])
            assert frac_z >= SYMBOL_MATCH_MIN, (t, "z symbols", frac_z, bad_z)
            for k in ("bpp", "bpp_y", "bpp_z"):
                assert rel_err(c[k].cpu(), o[k]) <= BPP_REL_TOL, (t, k)
            po, ro = gc.metrics(o["dpb"]["frame"], frames[:, t], masks[:, t])
            pc, rc = gc.metrics(c["dpb"]["frame"].cpu(), frames[:, t], masks[:, t])
            assert abs(po - pc) <= PSNR_TOL_DB and abs(ro - rc) <= PSNR_TOL_DB, (t, po, pc, ro, rc)
            if o.get("mask_pred") is not None:
                mo, mc = o["mask_pred"], c["mask_pred"].cpu()
                assert float((mo - mc).abs().max()) <= 1e-4 * max(1.0, float(mo.abs().max()))
            dpb_o = o["dpb"]
    del mp
    torch.cuda.empty_cache()


def test_recon_single_term_against_split_product():
    """recon_generation_net runs with plain fp16 operands (hi planes only) by default (x_hat of a P frame feeds no later
    symbol).  Against the fp32-grade product everywhere: identical symbols / bpp / feature, PSNR within
    the 0.02 dB gate with a wide margin, and a small element-wise x_hat difference."""
    case = gc.case_by_name("anchor_256")
    frames, masks = gc.case_inputs(case)
    outs = {}
    for name, flags in (("default", capi.FLAG_KEEP_TAPS), ("split3", capi.FLAG_KEEP_TAPS | capi.FLAG_RECON_SPLIT3)):
        _, mp = seeded_models("performance", case)
        mp = mp.cuda()
        mp.engine_flags = flags
        fr, mk = frames.cuda(), masks.cuda()
        with torch.no_grad():
            x_in = torch.cat([fr[:, 1], mk[:, 1]], 1)
            r = mp(x_in, 40, {"frame": fr[:, 0], "feature": None}, after_i=True)
            outs[name] = (r, mp.get_tap("y_q", x_in).cpu())
    (rd, yd), (rs, ys) = outs["default"], outs["split3"]
    assert torch.equal(yd, ys)
    assert torch.equal(rd["bpp"], rs["bpp"])
    assert torch.equal(rd["dpb"]["feature"], rs["dpb"]["feature"])
    xd, xs = rd["dpb"]["frame"].cpu(), rs["dpb"]["frame"].cpu()
    assert float((xd - xs).abs().max()) < 4e-3 and float((xd - xs).abs().mean()) < 4e-4
    pd, rod = gc.metrics(xd, frames[:, 1], masks[:, 1])
    ps, ros = gc.metrics(xs, frames[:, 1], masks[:, 1])
    assert abs(pd - ps) <= PSNR_TOL_DB / 4 and abs(rod - ros) <= PSNR_TOL_DB / 4
