"""GPU bring-up diagnostics (not a test): runs each section in its own process so that a trapped
kernel only kills that section, and prints per-stage differences against the oracle.

    python tests/diag/gpu_diag.py [section ...]      sections: ops_simt ops_umma gop_simt gop_umma
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TAPS = ["feature_in", "ctx", "ctx_t", "y_enc", "y", "hyper_in", "z", "z_hat", "hier", "temporal", "params",
        "y_hat_0", "spatial_prior", "y_q", "scales_hat", "y_hat"]


def ops(backend):
    import torch
    import torch.nn.functional as F
    import test_gpu_parity as T
    for cfg in T.CONV_CASES:
        cin, cout, k, s, p, H, W, act = cfg
        g = torch.Generator().manual_seed(cin * 7 + cout + k)
        x = torch.randn(2, cin, H, W, generator=g)
        w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        b = torch.randn(cout, generator=g)
        ref = T._act(F.conv2d(x, w, b, stride=s, padding=p), act)
        try:
            out = T.op_conv2d(x, w, b, s, p, 1, act, 3, backend)
            d = (out - ref).abs()
            print(f"conv {cfg}: max_err={float(d.max()):.3e} ref_max={float(ref.abs().max()):.3f} "
                  f"bad_frac={float((d > 1e-3).float().mean()):.4f}", flush=True)
        except Exception as e:   # noqa: BLE001
            print(f"conv {cfg}: EXC {e}", flush=True)
            break


def gop(backend, variant="old", case_name="anchor_256"):
    import torch
    import test_gpu_parity as T
    from helpers import gc, symbol_match
    case = gc.case_by_name(case_name)
    flags = T.capi.FLAG_KEEP_TAPS | (T.capi.FLAG_SIMT_GEMM if backend == 1 else 0)
    rep = T._run_case(variant, case, flags, record_taps=TAPS)
    for tag, o, c, to, tc, target, mask in rep:
        print(f"== {variant} {tag}: bpp oracle {o['bpp'].tolist()} cuda {c['bpp'].cpu().tolist()} "
              f"bpp_z {o['bpp_z'].tolist()} {c['bpp_z'].cpu().tolist()}")
        d = (o['dpb']['frame'] - c['dpb']['frame'].cpu()).abs()
        print(f"   x_hat max diff {float(d.max()):.3e} mean {float(d.mean()):.3e}")
        if o['dpb'].get('feature') is not None:
            d = (o['dpb']['feature'] - c['dpb']['feature'].cpu()).abs()
            print(f"   feature max diff {float(d.max()):.3e} ref max {float(o['dpb']['feature'].abs().max()):.3e}")
        for n in TAPS:
            if n in tc and n in to:
                a, b = to[n], tc[n]
                if a.shape != b.shape:
                    print(f"   tap {n}: SHAPE {tuple(a.shape)} vs {tuple(b.shape)}")
                    continue
                dd = (a - b).abs()
                extra = ""
                if n in ("y_q", "z_hat"):
                    extra = " match=%.6f bad=%d" % symbol_match(b, a)
                print(f"   tap {n}: max diff {float(dd.max()):.3e} ref max {float(a.abs().max()):.3e}{extra}")
    sys.stdout.flush()


def dcb(backend):
    import ctypes
    import torch
    import test_gpu_parity as T
    from helpers import D, O
    for cin, cout, shortcut, use_q in [(256, 256, False, False), (512, 256, False, True), (128, 128, True, False),
                                       (256, 320, False, True), (368, 368, False, False), (368, 192, False, False),
                                       (192, 368, False, True)]:
        g = torch.Generator().manual_seed(cin + cout)
        m = D.modules._dcb(cin, cout)
        sd = {"b." + k: v.detach() for k, v in m.state_dict().items()}
        x = torch.randn(2, cin, 16, 24, generator=g)
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
        out = torch.empty(2, cout, 16, 24, device="cuda")
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = T.capi.load().dmc_op_depth_conv_block(T._p(xb), ptrs, T._p(qb), T._p(out), 2, cin, cout, 16, 24,
                                                   int(shortcut), 3, backend, st)
        T.capi.check(rc, None)
        d = (out.cpu() - ref).abs()
        print(f"dcb {cin}->{cout} sc={shortcut} q={use_q}: max_err={float(d.max()):.3e} ref_max={float(ref.abs().max()):.3f} "
              f"bad_frac(>1e-4)={float((d > 1e-4).float().mean()):.5f}", flush=True)


def intra(backend, case_name="anchor_256"):
    import torch
    import test_gpu_parity as T
    from helpers import O, gc, sd_of, seeded_models, symbol_match
    case = gc.case_by_name(case_name)
    frames, masks = gc.case_inputs(case)
    mi, mp = seeded_models("old", case)
    sd_i, sd_p = sd_of(mi), sd_of(mp)
    mi, mp = mi.cuda(), mp.cuda()
    mi.engine_flags = mp.engine_flags = T.capi.FLAG_KEEP_TAPS | (T.capi.FLAG_SIMT_GEMM if backend == 1 else 0)
    with torch.no_grad():
        to = {}
        o = O.dmci_forward(sd_i, frames[:, 0], case["qp"], to)
        c = mi(frames[:, 0].cuda(), case["qp"])
        print("intra bpp", o["bpp"].tolist(), c["bpp"].cpu().tolist())
        for n in ["y", "z", "z_hat", "params", "y_q", "scales_hat", "y_hat"]:
            a, b = to[n], mi.get_tap(n, frames[:, 0].cuda()).cpu()
            dd = (a - b).abs()
            extra = " match=%.6f bad=%d" % symbol_match(b, a) if n in ("y_q", "z_hat") else ""
            print(f"   tap {n}: max diff {float(dd.max()):.3e} ref max {float(a.abs().max()):.3e}{extra}")
        d = (o["dpb"]["frame"] - c["dpb"]["frame"].cpu()).abs()
        print(f"   x_hat max diff {float(d.max()):.3e} mean {float(d.mean()):.3e}")
        # teacher-forced P frame: CUDA P model fed with the ORACLE's dpb
        for variant in ("old",):
            to = {}
            po = O.dmc_forward(sd_p, variant, frames[:, 1], 40, o["dpb"], True, to)
            pc = mp(frames[:, 1].cuda(), 40, {"frame": o["dpb"]["frame"].cuda(), "feature": None}, True)
            print("teacher-forced P1 bpp", po["bpp"].tolist(), pc["bpp"].cpu().tolist())
            for n in TAPS:
                if n in to:
                    a, b = to[n], mp.get_tap(n, frames[:, 1].cuda()).cpu()
                    dd = (a - b).abs()
                    extra = " match=%.6f bad=%d" % symbol_match(b, a) if n in ("y_q", "z_hat") else ""
                    print(f"   tap {n}: max diff {float(dd.max()):.3e} ref max {float(a.abs().max()):.3e}{extra}")
            d = (po["dpb"]["frame"] - pc["dpb"]["frame"].cpu()).abs()
            print(f"   x_hat max diff {float(d.max()):.3e} mean {float(d.mean()):.3e}")


SECTIONS = {"dcb_umma": lambda: dcb(0), "dcb_simt": lambda: dcb(1), "intra_umma": lambda: intra(0),"ops_simt": lambda: ops(1), "ops_umma": lambda: ops(0), "gop_simt": lambda: gop(1),
            "gop_umma": lambda: gop(0),
            "gop_umma_perf": lambda: gop(0, "performance", "rect_128x192"),
            "gop_umma_mp": lambda: gop(0, "mask_prop", "rect_128x192"),
            "gop_simt_mp": lambda: gop(1, "mask_prop", "rect_128x192")}

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--run":
        SECTIONS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(SECTIONS)
    for n in names:
        print(f"######## {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], timeout=600,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-6000:])
            print(f"######## {n} exit {r.returncode}", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"######## {n} TIMEOUT\n{(e.stdout or b'')[-3000:]}", flush=True)
