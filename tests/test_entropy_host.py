"""Range coder, host side (no GPU): the C oracle (oracle/rans_oracle.c) round-trips and stays near the ideal code length,
the product's table construction equals the reference's own Python (run from oracle/_ref with the absent native
pmf_to_quantized_cdf replaced by the oracle's), the container helpers are byte-compatible with stream_helper.py."""
import hashlib
import io

import numpy as np
import pytest
import torch

from helpers import D
from oracle import make_ref, rans_oracle as R

entropy = D.entropy
bitstream = D.bitstream


def _random_tables(rng, n_cdf=12, max_len=17):
    cdf = np.zeros((n_cdf, max_len + 2), dtype=np.int32)
    lens, offs = [], []
    for i in range(n_cdf):
        n = int(rng.integers(3, max_len + 1)) | 1
        pmf = rng.random(n) ** 3 + 1e-6
        pmf /= pmf.sum() * 1.02                     # leave ~2 % tail mass
        c = R.pmf_to_quantized_cdf(list(pmf) + [1.0 - pmf.sum()])
        cdf[i, : len(c)] = c
        lens.append(n + 2)
        offs.append(-(n // 2))
    return R.Tables(cdf, lens, offs)


def test_pmf_to_quantized_cdf_properties_and_agreement():
    rng = np.random.default_rng(0)
    for _ in range(200):
        n = int(rng.integers(2, 40))
        pmf = rng.random(n) ** 8                    # many near-empty symbols: exercises the stealing loop
        pmf /= pmf.sum()
        a = R.pmf_to_quantized_cdf(pmf)
        b = entropy.EntropyCoder.pmf_to_quantized_cdf(pmf.astype(np.float32).tolist()).numpy()
        assert np.array_equal(a, b)
        assert a[0] == 0 and a[-1] == 65536 and np.all(np.diff(a) > 0)


def test_oracle_round_trip_with_escapes_and_ragged_length():
    rng = np.random.default_rng(1)
    t = _random_tables(rng)
    for n in (0, 1, 255, 256, 257, 5000):
        idx = rng.integers(0, len(t.cdf_len), n).astype(np.int32)
        sym = np.round(rng.normal(0, 3, n)).astype(np.int32)
        if n > 10:
            sym[:6] = [300, -300, 2 ** 20, -(2 ** 20), 0, 9]          # far outside every table: bypass groups
        data = R.encode(t, sym, idx)
        assert np.array_equal(R.decode(t, data, idx), sym)
        streams = (n + 255) // 256
        assert len(data) <= 8 + 2 * streams + 4 * streams + R.ideal_bits(t, sym, idx) / 8 * 1.01 + 8 if n else len(data) == 8


def test_oracle_known_answer():
    """Pins the oracle's byte format (a change of the container or of the coder shows up here first)."""
    rng = np.random.default_rng(7)
    t = _random_tables(rng, n_cdf=5, max_len=9)
    idx = rng.integers(0, 5, 700).astype(np.int32)
    sym = np.round(rng.normal(0, 2.5, 700)).astype(np.int32)
    data = R.encode(t, sym, idx)
    assert len(data) == int(np.frombuffer(data[8:14], dtype="<u2").sum()) + 8 + 6
    assert hashlib.sha256(data).hexdigest()[:16] == KNOWN_DIGEST, hashlib.sha256(data).hexdigest()[:16]


KNOWN_DIGEST = "fdf6a4ff624f4f99"


def test_corrupt_container_is_rejected():
    rng = np.random.default_rng(2)
    t = _random_tables(rng)
    idx = rng.integers(0, len(t.cdf_len), 600).astype(np.int32)
    sym = np.round(rng.normal(0, 2, 600)).astype(np.int32)
    data = bytearray(R.encode(t, sym, idx))
    with pytest.raises(ValueError):
        R.decode(t, bytes(data[:-3]), idx)
    with pytest.raises(ValueError):
        R.decode(t, bytes(data), idx[:500])


def _reference_tables():
    """The reference's GaussianEncoder / BitEstimator table construction, run from oracle/_ref with its (absent) native
    pmf_to_quantized_cdf replaced by the oracle's."""
    make_ref.load_reference()
    from src.models import entropy_models as E

    class FakeCoder:
        def __init__(self):
            self.groups = []

        def add_cdf(self, cdf, cdf_length, offset):
            self.groups.append((np.array(cdf), np.array(cdf_length), np.array(offset)))
            return len(self.groups) - 1

    E.EntropyCoder.pmf_to_quantized_cdf = staticmethod(
        lambda pmf, precision=16: torch.IntTensor(R.pmf_to_quantized_cdf(pmf.tolist(), precision)))
    return E, FakeCoder


@pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_gaussian_tables_equal_the_reference_construction():
    E, FakeCoder = _reference_tables()
    ref = E.GaussianEncoder()
    fc = FakeCoder()
    ref.update(fc)
    cdf_r, len_r, off_r = fc.groups[0]
    mine = entropy.GaussianEncoder()
    cdf_m, len_m, off_m = mine.tables()
    assert torch.equal(mine.scale_table, ref.scale_table)
    assert np.array_equal(cdf_m.numpy(), cdf_r) and np.array_equal(len_m.numpy(), len_r) and np.array_equal(off_m.numpy(), off_r)
    assert cdf_m.shape[0] == 128 and int(len_m.max()) == cdf_m.shape[1]
    assert mine.log_step_recip == ref.log_step_recip


@pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_factorized_tables_equal_the_reference_construction():
    E, FakeCoder = _reference_tables()
    torch.manual_seed(3)
    ref = E.BitEstimator(6, 16)
    with torch.no_grad():
        for p in ref.parameters():
            p.mul_(40.0)                            # away from the near-uniform init ...
        for f in (ref.f1, ref.f2, ref.f3, ref.f4):
            f.h[:, ::2] += 1.2                      # ... and steeper cdfs on every other channel: different table lengths
    fc = FakeCoder()
    ref.update(fc)
    cdf_r, len_r, off_r = fc.groups[0]
    est = D.modules._bit_estimator(6, 16)
    est.load_state_dict(ref.state_dict())
    mine = entropy.BitEstimatorCoder(est, 6, 16)
    cdf_m, len_m, off_m = mine.tables()
    assert np.array_equal(cdf_m.numpy(), cdf_r) and np.array_equal(len_m.numpy(), len_r) and np.array_equal(off_m.numpy(), off_r)
    assert len(set(len_r.tolist())) > 1


def test_container_helpers():
    f = io.BytesIO()
    for v in (0, 127, 128, 16383, 16384, (1 << 30) - 1):
        bitstream.write_uint_adaptive(f, v)
    f.seek(0)
    assert [bitstream.read_uint_adaptive(f) for _ in range(6)] == [0, 127, 128, 16383, 16384, (1 << 30) - 1]
    with pytest.raises(ValueError):
        bitstream.write_uint_adaptive(io.BytesIO(), 1 << 30)
    f = io.BytesIO()
    sps = {"sps_id": 3, "height": 1280, "width": 1920, "ec_part": 1, "use_ada_i": 0}
    n = bitstream.write_sps(f, sps)
    n += bitstream.write_ip(f, False, 3, 40, bitstream.pack_streams(b"zz", b"y" * 300))
    assert n == len(f.getvalue())
    f.seek(0)
    h = bitstream.read_header(f)
    assert h["nal_type"] == bitstream.NalType.NAL_SPS and bitstream.read_sps_remaining(f, h["sps_id"]) == sps
    h = bitstream.read_header(f)
    qp, payload = bitstream.read_ip_remaining(f)
    assert h["nal_type"] == bitstream.NalType.NAL_P and qp == 40 and bitstream.unpack_streams(payload) == [b"zz", b"y" * 300]
    helper = bitstream.SPSHelper()
    assert helper.get_sps_id(sps) == (0, True) and helper.get_sps_id(sps) == (0, False)
    assert helper.get_sps_id(dict(sps, width=960)) == (1, True) and helper.get_sps_by_id(1)["width"] == 960


@pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_container_helpers_are_byte_compatible_with_the_reference():
    make_ref.load_reference()
    from src.utils import stream_helper as S
    sps = {"sps_id": 2, "height": 1280, "width": 1920, "ec_part": 0, "use_ada_i": 1}
    a, b = io.BytesIO(), io.BytesIO()
    assert S.write_sps(a, sps) == bitstream.write_sps(b, sps)
    for qp, payload in ((0, b""), (71, b"x" * 100), (255, b"y" * 20000)):
        assert S.write_ip(a, qp == 0, 2, qp, payload) == bitstream.write_ip(b, qp == 0, 2, qp, payload)
    assert a.getvalue() == b.getvalue()
    a.seek(0)
    hdr = S.read_header(a)
    assert S.read_sps_remaining(a, hdr["sps_id"]) == sps
    for v in (5, 200, 70000):
        x, y = io.BytesIO(), io.BytesIO()
        S.write_uint_adaptive(x, v)
        bitstream.write_uint_adaptive(y, v)
        assert x.getvalue() == y.getvalue()
