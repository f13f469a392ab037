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


@pytest.mark.parametrize("variant", ["old", "performance"])
def test_frame_symbols_round_trip_and_actual_bits(variant):
    """forward() -> the frame's y / z symbols coded by the GPU coder -> decoded bit-exactly; the container equals the
    oracle's; its size is the ideal code length of the tables (+ stream overhead); and -- reported, not gated -- how
    that compares with the bpp forward() estimates (the estimate clamps sigma at 1e-5, the coder's table starts at
    0.11: with random-init weights half the predicted sigmas are negative)."""
    H, W, qp = 256, 384, 32
    frames, masks = D.clips.synthetic_clip(11, 1, 2, H, W)
    torch.manual_seed(gc.SEED_P)
    m = D.build_p_model(variant).eval().cuda()
    m.engine_flags = D._capi.FLAG_KEEP_TAPS
    x = frames[:, 1].cuda() if variant == "old" else torch.cat([frames[:, 1], masks[:, 1]], 1).cuda()
    with torch.no_grad():
        r = m(x, qp, {"frame": frames[:, 0].cuda(), "feature": None}, after_i=True)
    fc = entropy.FrameCoder(m)
    streams = fc.compress(x, qp)
    y_q, scales, z_hat = m.get_tap("y_q", x), m.get_tap("scales_hat", x), m.get_tap("z_hat", x)
    y, z = fc.decompress_symbols(streams, scales, qp)
    assert torch.equal(y, y_q) and torch.equal(z, z_hat)
    # the same bytes from the CPU oracle
    ty, tz = _tables_of(fc.gaussian), R.Tables(*[a.numpy() for a in fc.z.tables()])
    iy = fc.gaussian.build_indexes(scales).cpu().numpy()
    iz = fc.z.build_indexes(z_hat.shape, qp, z_hat.device).cpu().numpy()
    assert streams["y"] == R.encode(ty, y_q.cpu().numpy().astype(np.int32), iy)
    assert streams["z"] == R.encode(tz, z_hat.cpu().numpy().astype(np.int32), iz)
    ideal = R.ideal_bits(ty, y_q.cpu().numpy(), iy) + R.ideal_bits(tz, z_hat.cpu().numpy(), iz)
    n_streams = (y_q.numel() + 255) // 256 + (z_hat.numel() + 255) // 256
    assert streams["bits"] <= 1.02 * ideal + 48 * n_streams + 128
    payload = D.bitstream.pack_streams(streams["z"], streams["y"])
    assert D.bitstream.unpack_streams(payload) == (streams["z"], streams["y"])
    actual_bpp = 8 * len(payload) / (H * W)
    print(f"{variant}: estimated bpp {float(r['bpp']):.4f} (sigma clamped at 1e-5), coded {actual_bpp:.4f} bpp "
          f"({len(payload)} bytes; ideal code length of the tables {ideal / (H * W):.4f} bpp)")
    assert math.isfinite(actual_bpp) and actual_bpp > 0
