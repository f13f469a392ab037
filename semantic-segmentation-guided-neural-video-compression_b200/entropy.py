"""Entropy coding of the quantised latents: the arithmetic-coding half of the reference's entropy models, on the GPU.

    EntropyCoder            src/models/entropy_models.py:11-81      (native RansEncoder / RansDecoder, absent upstream)
    GaussianEncoder         src/models/entropy_models.py:227-341    (y: 128 log-spaced scales, 16-bit cdfs)
    BitEstimator.update /
      encode_z / decode_z   src/models/entropy_models.py:152-224    (z: one factorized cdf per qp and channel)
    build_index_enc / _dec  src/layers/inference.py:63-84

Same names and argument meaning as the reference classes.  Table construction follows the reference's Python line by
line (it runs on whatever device the tensors are on); the coder itself is csrc/rans.cu behind the C ABI
(dmc_rans_* in include/dmc_b200.h).  There is no CPU coder in the product: without a CUDA device `update()` raises.

What this adds to forward(): actual bits and a decoder.  `FrameCoder.compress` codes the symbols of the last forward
(z_hat with the factorized tables, then y step by step with the scales each checkerboard step predicted) into the
frame payload; `FrameCoder.decompress` rebuilds x_hat / feature from the payload and the dpb alone, running the
decoder half of the network (dmc_decode_* in include/dmc_b200.h) with the range decoder between its phases.
"""
from __future__ import annotations

import ctypes
import math
from typing import Optional, Tuple

import torch

from . import _capi

__all__ = ["EntropyCoder", "GaussianEncoder", "BitEstimatorCoder", "FrameCoder", "owner_mask"]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class EntropyCoder:
    """Owner of the device-side cdf tables (the reference's EntropyCoder holds the native encoder / decoder objects)."""

    def __init__(self):
        self._lib = None
        self._groups = []          # (handle, device) per add_cdf call

    def _library(self):
        if self._lib is None:
            self._lib = _capi.load()
        return self._lib

    def __del__(self):
        try:
            for h, _ in self._groups:
                self._lib.dmc_rans_destroy(h)
        except Exception:
            pass

    @staticmethod
    def pmf_to_quantized_cdf(pmf, precision: int = 16) -> torch.Tensor:
        """CompressAI's pmf_to_quantized_cdf (what MLCodec_extensions_cpp exports under this name): round to
        `precision` bits, rescale so the total is exactly 2^precision, then steal one count from the cheapest symbol
        with more than one for every symbol that ended up empty."""
        p = [float(v) for v in pmf]
        cdf = [0] + [int(math.floor(v * (1 << precision) + 0.5)) for v in p]       # std::round on non-negative values
        total = sum(cdf)
        if total <= 0:
            raise ValueError("pmf_to_quantized_cdf: empty pmf")
        cdf = [((1 << precision) * c) // total for c in cdf]
        for i in range(1, len(cdf)):
            cdf[i] += cdf[i - 1]
        cdf[-1] = 1 << precision
        n = len(cdf) - 1
        for i in range(n):
            if cdf[i] == cdf[i + 1]:
                best_freq, best = None, -1
                for j in range(n):
                    f = cdf[j + 1] - cdf[j]
                    if f > 1 and (best_freq is None or f < best_freq):
                        best_freq, best = f, j
                if best < 0:
                    raise ValueError("pmf_to_quantized_cdf: cannot make every symbol codable")
                if best < i:
                    for j in range(best + 1, i + 1):
                        cdf[j] -= 1
                else:
                    for j in range(i + 1, best + 1):
                        cdf[j] += 1
        return torch.tensor(cdf, dtype=torch.int32)

    @staticmethod
    def pmf_to_cdf(pmf, tail_mass, pmf_length, max_length) -> torch.Tensor:
        """entropy_models.py:26-34: per table, the first pmf_length entries + the tail mass -> quantised cdf."""
        pmf, tail_mass = pmf.detach().float().cpu(), tail_mass.detach().float().cpu()
        lengths = [int(v) for v in pmf_length]
        cdf = torch.zeros((len(lengths), int(max_length) + 2), dtype=torch.int32)
        for i, n in enumerate(lengths):
            prob = torch.cat((pmf[i, :n], tail_mass[i].reshape(-1)[:1]), dim=0)
            c = EntropyCoder.pmf_to_quantized_cdf(prob.tolist(), 16)
            cdf[i, : c.numel()] = c
        return cdf

    def add_cdf(self, cdf, cdf_length, offset, device) -> int:
        """Uploads one group of tables; returns its cdf_group_index (entropy_models.py:39-43)."""
        lib = self._library()
        cdf = torch.as_tensor(cdf, dtype=torch.int32).contiguous().cpu()
        cdf_length = torch.as_tensor(cdf_length, dtype=torch.int32).reshape(-1).contiguous().cpu()
        offset = torch.as_tensor(offset, dtype=torch.int32).reshape(-1).contiguous().cpu()
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = lib.dmc_rans_create(_ptr(cdf), _ptr(cdf_length), _ptr(offset), cdf.shape[0], cdf.shape[1], ctypes.byref(h))
        if rc != 0:
            raise _capi.EngineError(f"dmc_rans_create failed ({rc}): {lib.dmc_rans_last_error(None).decode()}")
        self._groups.append((h.value, torch.device(device)))
        return len(self._groups) - 1

    # -- coding of flat device arrays ---------------------------------------------------------------
    def encode(self, group: int, symbols: torch.Tensor, indexes: torch.Tensor) -> bytes:
        lib = self._library()
        h, dev = self._groups[group]
        sym = symbols.detach().to(torch.float32).reshape(-1).contiguous()
        idx = indexes.reshape(-1).contiguous()
        assert sym.is_cuda and idx.is_cuda and idx.dtype == torch.int32 and sym.numel() == idx.numel()
        cap = int(lib.dmc_rans_max_bytes(sym.numel()))
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        nbytes = ctypes.c_int64()
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            rc = lib.dmc_rans_encode(h, _ptr(sym), _ptr(idx), sym.numel(), _ptr(out), cap, ctypes.byref(nbytes), st)
        if rc != 0:
            raise _capi.EngineError(f"dmc_rans_encode failed ({rc}): {lib.dmc_rans_last_error(h).decode()}")
        return bytes(out[: nbytes.value].cpu().numpy().tobytes())

    def decode(self, group: int, stream: bytes, indexes: torch.Tensor) -> torch.Tensor:
        lib = self._library()
        h, dev = self._groups[group]
        idx = indexes.reshape(-1).contiguous()
        buf = torch.frombuffer(bytearray(stream), dtype=torch.uint8).to(dev)
        out = torch.empty(idx.numel(), dtype=torch.float32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            rc = lib.dmc_rans_decode(h, _ptr(buf), buf.numel(), _ptr(idx), idx.numel(), _ptr(out), st)
        if rc != 0:
            raise _capi.EngineError(f"dmc_rans_decode failed ({rc}): {lib.dmc_rans_last_error(h).decode()}")
        return out


class GaussianEncoder:
    """entropy_models.py:227-341."""

    def __init__(self):
        self.scale_min = 0.11
        self.scale_max = 16.0
        self.scale_level = 128
        self.scale_table = self.get_scale_table(self.scale_min, self.scale_max, self.scale_level)
        self.log_scale_min = math.log(self.scale_min)
        self.log_scale_max = math.log(self.scale_max)
        self.log_scale_step = (self.log_scale_max - self.log_scale_min) / (self.scale_level - 1)
        self.log_step_recip = 1.0 / self.log_scale_step
        self.entropy_coder: Optional[EntropyCoder] = None
        self.cdf_group_index = None
        self._cdf_info = None

    @staticmethod
    def get_scale_table(min_val, max_val, levels):
        return torch.exp(torch.linspace(math.log(min_val), math.log(max_val), levels))

    def tables(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(quantized_cdf [128, max_length + 2], cdf_length, offset) -- entropy_models.py:247-280."""
        if self._cdf_info is None:
            scales = self.scale_table
            normal = torch.distributions.normal.Normal(0.0, scales)
            pmf_center = torch.full_like(scales, 8)
            for i in range(8, 1, -1):            # the smallest i in 2..8 whose cdf exceeds 0.9999 (8 if none does)
                probs = normal.cdf(torch.full_like(scales, float(i)))
                pmf_center = torch.where(probs > 0.9999, torch.full_like(scales, float(i)), pmf_center)
            pmf_center = pmf_center.int()
            pmf_length = 2 * pmf_center + 1
            max_length = int(pmf_length.max())
            samples = (torch.arange(max_length) - pmf_center[:, None]).float()
            normal2 = torch.distributions.normal.Normal(0.0, scales[:, None].expand_as(samples))
            upper = normal2.cdf(samples + 0.5)
            lower = normal2.cdf(samples - 0.5)
            pmf = upper - lower
            tail_mass = 2 * lower[:, :1]
            cdf = EntropyCoder.pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
            self._cdf_info = (cdf, (pmf_length + 2).int(), (-pmf_center).int())
        return self._cdf_info

    def update(self, entropy_coder: EntropyCoder, device, force_zero_thres=None):
        if force_zero_thres is not None:
            raise NotImplementedError("force_zero_thres (symbol skipping) is not part of this build")
        self.entropy_coder = entropy_coder
        self.cdf_group_index = entropy_coder.add_cdf(*self.tables(), device=device)

    def build_indexes(self, scales: torch.Tensor) -> torch.Tensor:
        """build_index_enc / build_index_dec (inference.py:63-84): clamp to [scale_min, scale_max], position in the
        log-spaced table (nearest entry).  The reference computes it in float and leaves the integer conversion to
        its native coder; negative / NaN predictions (the raw network output may be either) land on scale_min."""
        lib = self.entropy_coder._library()
        s = scales.detach().to(torch.float32).reshape(-1).contiguous()
        idx = torch.empty(s.numel(), dtype=torch.int32, device=s.device)
        st = ctypes.c_void_p(torch.cuda.current_stream(s.device).cuda_stream)
        with torch.cuda.device(s.device):
            rc = lib.dmc_rans_index_gaussian(_ptr(s), s.numel(), self.scale_min, self.scale_max, self.scale_level,
                                             _ptr(idx), st)
        if rc != 0:
            raise _capi.EngineError(f"dmc_rans_index_gaussian failed ({rc})")
        return idx

    def encode_y(self, x: torch.Tensor, scales: torch.Tensor) -> bytes:
        return self.entropy_coder.encode(self.cdf_group_index, x, self.build_indexes(scales))

    def decode_and_get_y(self, stream: bytes, scales: torch.Tensor, dtype, device) -> torch.Tensor:
        y = self.entropy_coder.decode(self.cdf_group_index, stream, self.build_indexes(scales.to(device)))
        return y.reshape(scales.shape).to(dtype)


class BitEstimatorCoder:
    """The coding half of BitEstimator (entropy_models.py:152-224) for a model's `bit_estimator_z` parameters."""

    def __init__(self, bit_estimator: torch.nn.Module, qp_num: int, channel: int):
        self.est = bit_estimator
        self.qp_num, self.channel = qp_num, channel
        self.entropy_coder: Optional[EntropyCoder] = None
        self.cdf_group_index = None

    def _cdf(self, x):
        """BitEstimator.get_cdf (entropy_models.py:139-150) for every (qp, channel): x is (qp_num, C, 1, L)."""
        for name in ("f1", "f2", "f3", "f4"):
            f = getattr(self.est, name)
            x = x * torch.nn.functional.softplus(f.h) + f.b
            if hasattr(f, "a"):
                x = x + torch.tanh(x) * torch.tanh(f.a)
        return torch.sigmoid(x)

    @torch.no_grad()
    def tables(self):
        """entropy_models.py:155-206."""
        dev = self.est.f1.h.device
        medians = torch.zeros((self.qp_num, self.channel, 1, 1), device=dev)
        minima = medians + 8
        for i in range(8, 1, -1):
            probs = self._cdf(torch.zeros_like(medians) - i)
            minima = torch.where(probs < 0.0001, torch.zeros_like(medians) + i, minima)
        maxima = medians + 8
        for i in range(8, 1, -1):
            probs = self._cdf(torch.zeros_like(medians) + i)
            maxima = torch.where(probs > 0.9999, torch.zeros_like(medians) + i, maxima)
        minima, maxima = minima.int(), maxima.int()
        offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length, device=dev)[None, None, None, :] + pmf_start
        lower = self._cdf(samples - 0.5)
        upper = self._cdf(samples + 0.5)
        pmf = (upper - lower)[:, :, 0, :]
        upper_max = self._cdf(maxima.to(torch.float32))
        tail_mass = lower[:, :, 0, :1] + (1.0 - upper_max[:, :, 0, -1:])
        pmf = pmf.reshape(-1, max_length)
        tail_mass = tail_mass.reshape(-1, 1)
        pmf_length = pmf_length.reshape(-1)
        cdf = EntropyCoder.pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        return cdf, (pmf_length + 2).int().cpu(), offset.reshape(-1).int().cpu()

    def update(self, entropy_coder: EntropyCoder):
        self.entropy_coder = entropy_coder
        self.cdf_group_index = entropy_coder.add_cdf(*self.tables(), device=self.est.f1.h.device)

    def build_indexes(self, size, qp: int, device) -> torch.Tensor:
        """entropy_models.py:208-211 for a flattened (B, C, H, W) tensor: qp * channel + c."""
        B, C, H, W = size
        lib = self.entropy_coder._library()
        n = B * C * H * W
        idx = torch.empty(n, dtype=torch.int32, device=device)
        st = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        with torch.cuda.device(device):
            rc = lib.dmc_rans_index_channels(n, H * W, C, int(qp) * self.channel, _ptr(idx), st)
        if rc != 0:
            raise _capi.EngineError(f"dmc_rans_index_channels failed ({rc})")
        return idx

    def encode_z(self, x: torch.Tensor, qp: int) -> bytes:
        return self.entropy_coder.encode(self.cdf_group_index, x, self.build_indexes(x.shape, qp, x.device))

    def decode_z(self, stream: bytes, size, qp: int, device, dtype=torch.float32) -> torch.Tensor:
        z = self.entropy_coder.decode(self.cdf_group_index, stream, self.build_indexes(size, qp, device))
        return z.reshape(size).to(dtype)


def owner_mask(step: int, shape, steps: int, device) -> torch.Tensor:
    """Boolean (B, C, H, W) mask of the latent elements coded in checkerboard step `step`
    (get_mask_2x / get_mask_4x, models/common_model.py:93-114,152-169; kernels.cu: prior_owner)."""
    B, C, H, W = shape
    c = torch.arange(C, device=device).view(1, C, 1, 1)
    h = torch.arange(H, device=device).view(1, 1, H, 1)
    w = torch.arange(W, device=device).view(1, 1, 1, W)
    if steps == 2:
        owner = (h + w + (c >= C // 2).long()) % 2
    else:
        table = torch.tensor([[0, 3, 2, 1], [3, 0, 1, 2], [2, 1, 0, 3], [1, 2, 3, 0]], device=device)
        owner = table[c // (C // 4), (h % 2) * 2 + (w % 2)]
    return (owner == step).expand(B, C, H, W)


class FrameCoder:
    """compress / decompress of one frame around a model of this package (P models and DMCI).

    compress()    after model(...) ran with FLAG_KEEP_TAPS: z_hat with the factorized tables, then the y symbols of
                  every checkerboard step with the scales that step predicted -- the order a decoder can follow
                  (the reference's dead compress(), src/models/video_model.py:256-296, codes y_q_w_0 / y_q_w_1 the
                  same way).  Returns the frame payload (bitstream.pack_streams) and its parts.
    decompress()  bytes + dpb + qp -> the decoder half of the network through dmc_decode_* with the range decoder
                  between the phases; returns {"dpb": {"frame", "feature"}} bit-identical to forward()'s."""

    def __init__(self, model):
        self.model = model
        dev = next(model.parameters()).device
        self.intra = model.variant == "intra"
        self.steps = 4 if self.intra else 2
        self.coder = EntropyCoder()
        self.gaussian = GaussianEncoder()
        self.gaussian.update(self.coder, dev)
        qp_num = model.bit_estimator_z.f1.h.shape[0]
        self.z = BitEstimatorCoder(model.bit_estimator_z, qp_num, model.bit_estimator_z.f1.h.shape[1])
        self.z.update(self.coder)

    @torch.no_grad()
    def compress(self, x_like: torch.Tensor, qp: int) -> dict:
        """x_like: the input of the forward that just ran (for the tap shapes)."""
        from . import bitstream
        m = self.model
        y_q = m.get_tap("y_q", x_like)
        scales = m.get_tap("scales_hat", x_like)
        z_hat = m.get_tap("z_hat", x_like)
        parts = [self.z.encode_z(z_hat, qp)]
        for k in range(self.steps):
            own = owner_mask(k, y_q.shape, self.steps, y_q.device)
            parts.append(self.gaussian.encode_y(y_q[own], scales[own]))
        payload = bitstream.pack_streams(*parts)
        return {"payload": payload, "z": parts[0], "y": parts[1:], "y_shape": tuple(y_q.shape),
                "z_shape": tuple(z_hat.shape), "bits": 8 * len(payload)}

    @torch.no_grad()
    def decompress(self, payload: bytes, shape, qp: int, dpb: Optional[dict] = None, after_i: bool = True) -> dict:
        """shape: (B, 3, H, W) of the frame.  dpb: {"frame", "feature"} for P frames (None for the intra model)."""
        from . import bitstream
        m = self.model
        B, _, H, W = shape
        dev = next(m.parameters()).device
        parts = bitstream.unpack_streams(payload)
        if len(parts) != 1 + self.steps:
            raise ValueError(f"frame payload holds {len(parts)} streams, expected {1 + self.steps}")
        h, stream = m._engine(B, H, W, dev)
        lib = m._lib
        Cy = 256 if self.intra else m.cfg.ch_y
        H16, W16 = H // 16, W // 16
        Hz, Wz = (H16 + 3) // 4, (W16 + 3) // 4
        z_shape = (B, self.z.channel, Hz, Wz)
        z_hat = self.z.decode_z(parts[0], z_shape, qp, dev).contiguous()
        frame = feature = None
        if not self.intra:
            if after_i:
                frame = dpb["frame"].contiguous()
                m._check_tensor("dpb['frame']", frame, (B, 3, H, W), dev)
            else:
                feature = dpb["feature"].contiguous()
                m._check_tensor("dpb['feature']", feature, (B, m.cfg.ch_d, H // 8, W // 8), dev)
        with torch.cuda.device(dev):
            _capi.check(lib.dmc_decode_begin(h, _ptr(frame) if frame is not None else None,
                                             _ptr(feature) if feature is not None else None, int(qp),
                                             1 if after_i else 0, _ptr(z_hat), stream), h)
            y_shape = (B, Cy, H16, W16)
            sigma = torch.empty(y_shape, dtype=torch.float32, device=dev)
            for k in range(self.steps):
                _capi.check(lib.dmc_decode_sigma(h, k, _ptr(sigma), stream), h)
                own = owner_mask(k, y_shape, self.steps, dev)
                sym = self.gaussian.decode_and_get_y(parts[1 + k], sigma[own], torch.float32, dev)
                dense = torch.zeros(y_shape, dtype=torch.float32, device=dev)
                dense[own] = sym
                _capi.check(lib.dmc_decode_symbols(h, k, _ptr(dense), stream), h)
            x_hat = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
            feat = None if self.intra else torch.empty((B, m.cfg.ch_d, H // 8, W // 8), dtype=torch.float32, device=dev)
            _capi.check(lib.dmc_decode_finish(h, _ptr(x_hat), _ptr(feat) if feat is not None else None, stream), h)
        return {"dpb": {"frame": x_hat, "feature": feat}}
