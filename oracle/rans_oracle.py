"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/rans_oracle.c (the CPU statement of the range coder; see the
header of that file for what it restates and why parity of the BIT FORMAT is unpinned: the reference's native coder,
MLCodec_extensions_cpp, is absent from its tree).  Only tests/ and __graft_entry__ use this module."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rans_oracle.c")
LIB = os.path.join(HERE, "_build", "librans_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -shared oracle/rans_oracle.c -> oracle/_build/librans_oracle.so (git-ignored, travels with gpurun)."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        P = ctypes.c_void_p
        lib.rans_oracle_pmf_to_cdf.restype = ctypes.c_int
        lib.rans_oracle_pmf_to_cdf.argtypes = [P, ctypes.c_int, ctypes.c_int, P]
        lib.rans_oracle_encode.restype = ctypes.c_int64
        lib.rans_oracle_encode.argtypes = [P, P, P, ctypes.c_int, ctypes.c_int, P, P, ctypes.c_int64, P, ctypes.c_int64]
        lib.rans_oracle_decode.restype = ctypes.c_int
        lib.rans_oracle_decode.argtypes = [P, P, P, ctypes.c_int, ctypes.c_int, P, ctypes.c_int64, P, ctypes.c_int64, P]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def pmf_to_quantized_cdf(pmf, precision: int = 16) -> np.ndarray:
    pmf = np.ascontiguousarray(pmf, dtype=np.float32)
    cdf = np.zeros(len(pmf) + 1, dtype=np.uint32)
    rc = _load().rans_oracle_pmf_to_cdf(_p(pmf), len(pmf), precision, _p(cdf))
    if rc:
        raise ValueError(f"pmf_to_quantized_cdf failed ({rc})")
    return cdf.astype(np.int32)


class Tables:
    def __init__(self, cdf, cdf_len, offset):
        self.cdf = np.ascontiguousarray(cdf, dtype=np.int32)
        self.cdf_len = np.ascontiguousarray(cdf_len, dtype=np.int32).reshape(-1)
        self.offset = np.ascontiguousarray(offset, dtype=np.int32).reshape(-1)
        assert self.cdf.ndim == 2 and len(self.cdf_len) == len(self.offset) == self.cdf.shape[0]


def encode(t: Tables, sym, idx) -> bytes:
    sym = np.ascontiguousarray(sym, dtype=np.int32).reshape(-1)
    idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1)
    n = len(sym)
    streams = (n + 255) // 256
    out = np.zeros(8 + 2 * streams + streams * 2560, dtype=np.uint8)
    size = _load().rans_oracle_encode(_p(t.cdf), _p(t.cdf_len), _p(t.offset), t.cdf.shape[0], t.cdf.shape[1], _p(sym),
                                      _p(idx), n, _p(out), len(out))
    if size < 0:
        raise ValueError("rans_oracle_encode: bad index or empty table entry")
    return out[:size].tobytes()


def decode(t: Tables, data: bytes, idx) -> np.ndarray:
    idx = np.ascontiguousarray(idx, dtype=np.int32).reshape(-1)
    buf = np.frombuffer(data, dtype=np.uint8)
    sym = np.zeros(len(idx), dtype=np.int32)
    rc = _load().rans_oracle_decode(_p(t.cdf), _p(t.cdf_len), _p(t.offset), t.cdf.shape[0], t.cdf.shape[1], _p(buf),
                                    len(buf), _p(idx), len(idx), _p(sym))
    if rc:
        raise ValueError(f"rans_oracle_decode: malformed container ({rc})")
    return sym


def ideal_bits(t: Tables, sym, idx) -> float:
    """Code length of in-range symbols under the tables: sum -log2(freq / 65536) (escapes counted as the escape entry
    plus 4 bits per bypass group incl. the count group)."""
    sym = np.asarray(sym, dtype=np.int64).reshape(-1)
    idx = np.asarray(idx, dtype=np.int64).reshape(-1)
    v = sym - t.offset[idx]
    mx = t.cdf_len[idx] - 2
    esc = (v < 0) | (v >= mx)
    raw = np.where(v < 0, -2 * v - 1, 2 * (v - mx))
    vv = np.where(esc, mx, v)
    freq = t.cdf[idx, vv + 1] - t.cdf[idx, vv]
    bits = -np.log2(freq / 65536.0)
    groups = np.where(esc, np.ceil(np.log2(np.maximum(raw, 1) + 1) / 4.0) + 1, 0)
    return float(bits.sum() + 4.0 * groups.sum())
