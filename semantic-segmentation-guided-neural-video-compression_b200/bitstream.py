"""The DCVC bit-stream container the reference carries in src/utils/stream_helper.py:68-217: a sequence of NAL-like
units, each starting with one byte `type << 4 | sps_id`.

    SPS   type 0: height, width (adaptive-length integers), one flag byte `ec_part << 2 | use_ada_i`
    I     type 1: one byte qp, adaptive-length payload size, payload
    P     type 2: same as I

Adaptive-length integers: 1 byte below 2^7 (top bit 0), 2 bytes below 2^14 (top bits 10), else 4 bytes below 2^30 (top
bits 11), big endian.  The payload of an I / P unit here is `pack_streams(...)`: the range-coder containers of a
frame in decoding order (z, then y of every checkerboard step), each preceded by its adaptive-length size.  Byte-compatible with the reference's helpers for the unit
headers (tests/test_entropy_host.py checks it against oracle/_ref when that copy is present).
"""
from __future__ import annotations

import enum
import io
from typing import BinaryIO, Dict, List, Tuple


class NalType(enum.IntEnum):
    NAL_SPS = 0
    NAL_I = 1
    NAL_P = 2


def write_uint_adaptive(f: BinaryIO, a: int) -> int:
    if a < 0 or a >= 1 << 30:
        raise ValueError("adaptive-length integers cover [0, 2^30)")
    if a < 1 << 7:
        f.write(bytes([a]))
        return 1
    if a < 1 << 14:
        f.write(bytes([0x80 | (a >> 8), a & 0xFF]))
        return 2
    f.write(bytes([0xC0 | (a >> 24), (a >> 16) & 0xFF, (a >> 8) & 0xFF, a & 0xFF]))
    return 4


def read_uint_adaptive(f: BinaryIO) -> int:
    b0 = f.read(1)[0]
    if b0 < 0x80:
        return b0
    b1 = f.read(1)[0]
    if b0 >> 6 == 0b10:
        return ((b0 & 0x3F) << 8) | b1
    b2, b3 = f.read(2)
    return ((b0 & 0x3F) << 24) | (b1 << 16) | (b2 << 8) | b3


def write_sps(f: BinaryIO, sps: Dict[str, int]) -> int:
    if not (0 <= sps["sps_id"] < 16 and sps["use_ada_i"] in (0, 1)):
        raise ValueError("sps_id must be below 16 and use_ada_i a flag")
    f.write(bytes([(int(NalType.NAL_SPS) << 4) | sps["sps_id"]]))
    n = 1 + write_uint_adaptive(f, sps["height"]) + write_uint_adaptive(f, sps["width"])
    f.write(bytes([(sps["ec_part"] << 2) | sps["use_ada_i"]]))
    return n + 1


def read_header(f: BinaryIO) -> Dict[str, int]:
    flag = f.read(1)[0]
    return {"nal_type": NalType(flag >> 4), "sps_id": flag & 0x0F}


def read_sps_remaining(f: BinaryIO, sps_id: int) -> Dict[str, int]:
    height = read_uint_adaptive(f)
    width = read_uint_adaptive(f)
    flag = f.read(1)[0]
    return {"sps_id": sps_id, "height": height, "width": width, "ec_part": (flag >> 2) & 1, "use_ada_i": flag & 1}


def write_ip(f: BinaryIO, is_i_frame: bool, sps_id: int, qp: int, bit_stream: bytes) -> int:
    if not 0 <= qp < 256:
        raise ValueError("qp must fit one byte")
    f.write(bytes([(int(NalType.NAL_I if is_i_frame else NalType.NAL_P) << 4) | sps_id, qp]))
    n = 2 + write_uint_adaptive(f, len(bit_stream))
    f.write(bit_stream)
    return n + len(bit_stream)


def read_ip_remaining(f: BinaryIO) -> Tuple[int, bytes]:
    qp = f.read(1)[0]
    size = read_uint_adaptive(f)
    return qp, f.read(size)


class SPSHelper:
    """Hands out sps ids for (height, width, use_ada_i, ec_part) combinations, at most 16 (stream_helper.py:108-137)."""

    def __init__(self):
        self.spss: List[Dict[str, int]] = []

    def get_sps_id(self, target: Dict[str, int]) -> Tuple[int, bool]:
        for sps in self.spss:
            if all(sps[k] == target[k] for k in ("height", "width", "use_ada_i", "ec_part")):
                return sps["sps_id"], False
        new_id = max((s["sps_id"] for s in self.spss), default=-1) + 1
        if new_id > 15:
            raise ValueError("more than 16 parameter sets")
        self.spss.append(dict(target, sps_id=new_id))
        return new_id, True

    def add_sps_by_id(self, sps: Dict[str, int]):
        self.spss = [s for s in self.spss if s["sps_id"] != sps["sps_id"]] + [dict(sps)]

    def get_sps_by_id(self, sps_id: int):
        return next((s for s in self.spss if s["sps_id"] == sps_id), None)


def pack_streams(*streams: bytes) -> bytes:
    """Payload of one frame: the range-coder containers in decoding order (z, then y step by step), each behind its
    adaptive-length size, with the number of containers in front."""
    f = io.BytesIO()
    write_uint_adaptive(f, len(streams))
    for s in streams:
        write_uint_adaptive(f, len(s))
        f.write(s)
    return f.getvalue()


def unpack_streams(payload: bytes) -> List[bytes]:
    f = io.BytesIO(payload)
    return [f.read(read_uint_adaptive(f)) for _ in range(read_uint_adaptive(f))]
