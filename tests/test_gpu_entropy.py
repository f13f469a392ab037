"""GPU range coder (csrc/rans.cu) against the C oracle (oracle/rans_oracle.c): byte-identical containers, exact round
trips, and actual coded size of a frame's symbols next to the estimate forward() reports."""
import math

import numpy as np
import pytest
import torch

from helpers import D, gc
from oracle import rans_oracle as R

pytestmark = pytest.mark.gpu
entropy = D.entropy


def _tables_of(enc):
    cdf, cdf_len, off = enc.tables()
    return R.Tables(cdf.numpy(), cdf_len.numpy(), off.numpy())


def _gaussian():
    coder = entropy.EntropyCoder()
    g = entropy.GaussianEncoder()
    g.update(coder, torch.device("cuda"))
    return coder, g


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 100_000])
def test_gaussian_streams_are_byte_identical_to_the_oracle(n):
    coder, g = _gaussian()
    t = _tables_of(g)
    gen = torch.Generator().manual_seed(n)
    sigma = torch.exp(torch.randn(n, generator=gen) * 1.5)          # 0.01 ... 100: clamped at both ends of the table
    if n > 16:
        sigma[:6] = torch.tensor([float("nan"), -3.0, 0.0, 1e-7, 1e6, 0.11])
    sym = torch.round(torch.randn(n, generator=gen) * sigma.nan_to_num(1.0).clamp(0.05, 40.0))
    if n > 16:
        sym[6:12] = torch.tensor([500.0, -500.0, 2.0 ** 20, -(2.0 ** 20), 9.0, -9.0])    # escapes with 1 ... 6 bypass groups
    idx = g.build_indexes(sigma.cuda())
    # the index kernel against the formula of build_index_enc (inference.py:76-84), nearest table entry
    s = sigma.nan_to_num(g.scale_min).clamp(g.scale_min, g.scale_max)
    want = torch.round((torch.log(s) - g.log_scale_min) * g.log_step_recip).clamp(0, 127).int()
    off = (idx.cpu() - want).abs()
    assert int(off.max()) <= 1 and float((off > 0).float().mean()) <= 1e-3 if n else True     # (ties of logf vs torch.log)
    data = coder.encode(g.cdf_group_index, sym.cuda(), idx)
    ref = R.encode(t, sym.numpy().astype(np.int32), idx.cpu().numpy())
    assert data == ref
    back = coder.decode(g.cdf_group_index, ref, idx).cpu()
    assert torch.equal(back, sym)
    assert np.array_equal(R.decode(t, data, idx.cpu().numpy()), sym.numpy().astype(np.int32))
    if n >= 1000:
        assert 8 * len(data) <= 1.02 * R.ideal_bits(t, sym.numpy(), idx.cpu().numpy()) + 48 * ((n + 255) // 256) + 64


def test_corrupt_or_mismatched_containers_raise():
    coder, g = _gaussian()
    sigma = torch.rand(3000).cuda() * 3 + 0.2
    sym = torch.round(torch.randn(3000).cuda() * sigma)
    idx = g.build_indexes(sigma)
    data = coder.encode(g.cdf_group_index, sym, idx)
    with pytest.raises(D._capi.EngineError):
        coder.decode(g.cdf_group_index, data[:-5], idx)
    with pytest.raises(D._capi.EngineError):
        coder.decode(g.cdf_group_index, data, idx[:2000])
    assert torch.equal(coder.decode(g.cdf_group_index, data, idx), sym)       # the coder is still usable


def _frame_inputs(variant, H, W, seed=11):
    frames, masks = D.clips.synthetic_clip(seed, 1, 3, H, W)

    def x_of(t):
        return frames[:, t].cuda() if variant == "old" else torch.cat([frames[:, t], masks[:, t]], 1).cuda()
    return frames, x_of


@pytest.mark.parametrize("variant", ["old", "performance"])
def test_frame_symbols_round_trip_and_actual_bits(variant):
    """forward() -> the frame's z / y symbols coded step by step by the GPU coder; every container equals the
    oracle's; its size is the ideal code length of the tables (+ stream overhead); and -- reported, not gated -- how
    that compares with the bpp forward() estimates (the estimate clamps sigma at 1e-5, the coder's table starts at
    0.11: with random-init weights half the predicted sigmas are negative)."""
    H, W, qp = 256, 384, 32
    frames, x_of = _frame_inputs(variant, H, W)
    torch.manual_seed(gc.SEED_P)
    m = D.build_p_model(variant).eval().cuda()
    m.engine_flags = D._capi.FLAG_KEEP_TAPS
    x = x_of(1)
    with torch.no_grad():
        r = m(x, qp, {"frame": frames[:, 0].cuda(), "feature": None}, after_i=True)
    fc = entropy.FrameCoder(m)
    streams = fc.compress(x, qp)
    y_q, scales, z_hat = m.get_tap("y_q", x), m.get_tap("scales_hat", x), m.get_tap("z_hat", x)
    # the same bytes from the CPU oracle, and the symbols back from them
    ty, tz = _tables_of(fc.gaussian), R.Tables(*[a.numpy() for a in fc.z.tables()])
    iz = fc.z.build_indexes(z_hat.shape, qp, z_hat.device).cpu().numpy()
    assert streams["z"] == R.encode(tz, z_hat.cpu().numpy().astype(np.int32), iz)
    assert torch.equal(fc.z.decode_z(streams["z"], tuple(z_hat.shape), qp, z_hat.device), z_hat)
    ideal = R.ideal_bits(tz, z_hat.cpu().numpy(), iz)
    n_streams = (z_hat.numel() + 255) // 256
    assert len(streams["y"]) == 2
    for k, part in enumerate(streams["y"]):
        own = entropy.owner_mask(k, y_q.shape, 2, y_q.device)
        sym, sg = y_q[own], scales[own]
        iy = fc.gaussian.build_indexes(sg).cpu().numpy()
        assert part == R.encode(ty, sym.cpu().numpy().astype(np.int32), iy)
        assert torch.equal(fc.gaussian.decode_and_get_y(part, sg, torch.float32, sg.device), sym)
        ideal += R.ideal_bits(ty, sym.cpu().numpy(), iy)
        n_streams += (sym.numel() + 255) // 256
    assert streams["bits"] <= 1.02 * ideal + 48 * n_streams + 256
    assert D.bitstream.unpack_streams(streams["payload"]) == [streams["z"], *streams["y"]]
    actual_bpp = streams["bits"] / (H * W)
    print(f"{variant}: estimated bpp {float(r['bpp']):.4f} (sigma clamped at 1e-5), coded {actual_bpp:.4f} bpp "
          f"({len(streams['payload'])} bytes; ideal code length of the tables {ideal / (H * W):.4f} bpp)")
    assert math.isfinite(actual_bpp) and actual_bpp > 0


@pytest.mark.parametrize("variant,H,W,B", [("old", 80, 112, 1), ("fast", 80, 112, 2), ("mask_prop", 144, 208, 1)])
def test_ragged_and_batched_frames_decode_from_bytes_alone(variant, H, W, B):
    """Sizes whose latent (H/16 x W/16) is not a multiple of 4: y is replicate-padded for the hyper path and the
    hyper-decoder output cropped back (models/common_model.py:68-72) -- on both sides of the split; and a batch of 2."""
    qp = 27
    frames, masks = D.clips.synthetic_clip(17, B, 3, H, W)
    torch.manual_seed(gc.SEED_P)
    m = D.build_p_model(variant).eval().cuda()
    m.engine_flags = D._capi.FLAG_KEEP_TAPS
    fc = entropy.FrameCoder(m)
    dpb_enc = {"frame": frames[:, 0].cuda(), "feature": None}
    dpb_dec = {"frame": frames[:, 0].cuda(), "feature": None}
    for t in (1, 2):
        x = frames[:, t].cuda() if variant == "old" else torch.cat([frames[:, t], masks[:, t]], 1).cuda()
        with torch.no_grad():
            r = m(x, qp, dpb_enc, after_i=(t == 1))
        want_frame, want_feat = r["dpb"]["frame"].clone(), r["dpb"]["feature"].clone()
        s = fc.compress(x, qp)
        assert s["z_shape"][2:] == ((H // 16 + 3) // 4, (W // 16 + 3) // 4)
        d = fc.decompress(s["payload"], (B, 3, H, W), qp, dpb_dec, after_i=(t == 1))
        assert torch.equal(d["dpb"]["frame"], want_frame) and torch.equal(d["dpb"]["feature"], want_feat)
        dpb_enc, dpb_dec = r["dpb"], d["dpb"]


@pytest.mark.parametrize("variant", ["old", "performance", "fast", "mask_prop"])
def test_p_frames_decode_from_bytes_alone(variant):
    """The decoder half (dmc_decode_* + the range decoder between its phases) rebuilds x_hat and the feature of both
    kinds of P frame from the payload, the dpb and qp -- bit-identical to what the encoder's forward() returned, so a
    decoder's dpb never drifts from the encoder's (video_model.py:256-333: the split the reference sketches)."""
    H, W, qp = 128, 192, 30
    frames, x_of = _frame_inputs(variant, H, W, seed=5)
    torch.manual_seed(gc.SEED_P)
    m = D.build_p_model(variant).eval().cuda()
    m.engine_flags = D._capi.FLAG_KEEP_TAPS
    fc = entropy.FrameCoder(m)
    dpb_enc = {"frame": frames[:, 0].cuda(), "feature": None}
    dpb_dec = {"frame": frames[:, 0].cuda(), "feature": None}
    for t in (1, 2):
        x = x_of(t)
        with torch.no_grad():
            r = m(x, qp, dpb_enc, after_i=(t == 1))
        payload = fc.compress(x, qp)["payload"]
        want_frame, want_feat = r["dpb"]["frame"].clone(), r["dpb"]["feature"].clone()
        d = fc.decompress(payload, (1, 3, H, W), qp, dpb_dec, after_i=(t == 1))
        assert torch.equal(d["dpb"]["frame"], want_frame), f"{variant} P{t}: x_hat differs"
        assert torch.equal(d["dpb"]["feature"], want_feat), f"{variant} P{t}: feature differs"
        dpb_enc, dpb_dec = r["dpb"], d["dpb"]
    with pytest.raises(ValueError):
        fc.decompress(D.bitstream.pack_streams(b"", b""), (1, 3, H, W), qp, dpb_dec, after_i=False)


def test_intra_frame_decodes_from_bytes_alone():
    H, W, qp = 128, 192, 37
    frames, _ = D.clips.synthetic_clip(3, 1, 1, H, W)
    torch.manual_seed(gc.SEED_I)
    m = D.DMCI().eval().cuda()
    m.engine_flags = D._capi.FLAG_KEEP_TAPS
    fc = entropy.FrameCoder(m)
    x = frames[:, 0].cuda()
    with torch.no_grad():
        r = m(x, qp)
    want = r["dpb"]["frame"].clone()
    s = fc.compress(x, qp)
    assert len(s["y"]) == 4
    d = fc.decompress(s["payload"], (1, 3, H, W), qp)
    assert torch.equal(d["dpb"]["frame"], want)
    print(f"intra: estimated bpp {float(r['bpp']):.4f}, coded {s['bits'] / (H * W):.4f} bpp")


def test_full_size_gop_decodes_from_bytes_alone():
    """BASELINE config 2 size (1920x1280, `performance`): intra frame + two P frames coded to bytes; a decoder that only
    ever sees the payloads (its dpb is what IT decoded) ends with frames bit-identical to the encoder's."""
    import time
    H, W, qp = 1280, 1920, 32
    frames, masks = D.clips.synthetic_clip(21, 1, 3, H, W)
    torch.manual_seed(gc.SEED_I)
    mi = D.DMCI().eval().cuda()
    torch.manual_seed(gc.SEED_P)
    mp = D.build_p_model("performance").eval().cuda()
    mi.engine_flags = mp.engine_flags = D._capi.FLAG_KEEP_TAPS
    ci, cp = entropy.FrameCoder(mi), entropy.FrameCoder(mp)
    x0 = frames[:, 0].cuda()
    with torch.no_grad():
        r = mi(x0, qp)
    s = ci.compress(x0, qp)
    d = ci.decompress(s["payload"], (1, 3, H, W), qp)
    assert torch.equal(d["dpb"]["frame"], r["dpb"]["frame"])
    print(f"I: estimated {float(r['bpp']):.4f} bpp, coded {s['bits'] / (H * W):.4f} bpp")
    dpb_enc, dpb_dec = r["dpb"], d["dpb"]
    for t in (1, 2):
        x = torch.cat([frames[:, t], masks[:, t]], 1).cuda()
        q = mp.shift_qp(qp, D.clips.INDEX_MAP[t % 8])
        with torch.no_grad():
            r = mp(x, q, dpb_enc, after_i=(t == 1))
        want_frame, want_feat = r["dpb"]["frame"].clone(), r["dpb"]["feature"].clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s = cp.compress(x, q)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        d = cp.decompress(s["payload"], (1, 3, H, W), q, dpb_dec, after_i=(t == 1))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        assert torch.equal(d["dpb"]["frame"], want_frame) and torch.equal(d["dpb"]["feature"], want_feat)
        print(f"P{t}: estimated {float(r['bpp']):.4f} bpp, coded {s['bits'] / (H * W):.4f} bpp ({len(s['payload'])} bytes); "
              f"compress {1e3 * (t1 - t0):.1f} ms, decompress {1e3 * (t2 - t1):.1f} ms")
        dpb_enc, dpb_dec = r["dpb"], d["dpb"]
