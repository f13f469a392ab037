"""Drop-in `nn.Module`s for the reference's codec models, backed by the CUDA engine.

    DMC_old / DMC_performance / DMC_fast / DMC_mask_prop     <->  the four `dmc_variant`s
    DMCI                                                      <->  the intra model

Same constructors, same parameter tree (state_dict keys, shapes, registration order and
default initialisation -- so `torch.manual_seed(s); Model()` gives the reference's weights),
same `forward(x, qp, dpb, after_i)` / `shift_qp` signatures and result dict as

    old          src/models/video_model.py:183-388
    performance  src/refactor/seg_video_model.py:205-365
    fast         src/refactor/seg_video_model_fast.py:185-411
    mask_prop    src/refactor/mask_prop_seg_video_model.py:185-417
    DMCI         src/models/image_model.py:96-261

The modules only hold parameters; all arithmetic happens in libdmc_b200.so through the C ABI
(include/dmc_b200.h).  There is no PyTorch or CPU fallback: calling forward without CUDA, in
training mode, or with autograd enabled raises.
"""
from __future__ import annotations

import collections
import ctypes
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _capi

__all__ = ["DMCConfig", "DMC_old", "DMC_performance", "DMC_fast", "DMC_mask_prop", "DMCI",
           "P_MODELS", "build_p_model", "NonFiniteError"]


@dataclass
class DMCConfig:
    """src/refactor/config.py:15-26."""
    patch_size: int = 8
    src: int = 3 * 8 * 8
    ch_d: int = 256
    ch_y: int = 128
    ch_z: int = 128
    ch_recon: int = 320
    qp_shift: Tuple[int, int, int] = (0, 8, 4)
    extra_qp: int = 8


# --------------------------------------------------------------------------
# parameter tree builders (containers only -- they have no forward)
# --------------------------------------------------------------------------
class _Tree(nn.Module):
    """Named container; children given as (name, module) in registration order."""

    def __init__(self, *children):
        super().__init__()
        for name, mod in children:
            self.add_module(str(name), mod)

    def __getitem__(self, name):
        """tree[3] / tree["down"]: the child registered under that name (like nn.Sequential indexing)."""
        return self._modules[str(name)]

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the computation lives in the CUDA engine")


def _c(cin, cout, k=1, stride=1, padding=0, groups=1):
    return nn.Conv2d(cin, cout, k, stride=stride, padding=padding, groups=groups)


def _dcb(cin, cout, force_adaptor=False):
    """Parameter layout of DepthConvBlock (layers.py:43-63)."""
    parts = []
    if cin != cout or force_adaptor:
        parts.append(("adaptor", _c(cin, cout)))
    parts.append(("dc", _Tree((0, _c(cout, cout)), (2, _c(cout, cout, 3, padding=1, groups=cout)),
                              (3, _c(cout, cout)))))
    parts.append(("ffn", _Tree((0, _c(cout, cout * 4)), (2, _c(cout * 2, cout)))))
    return _Tree(*parts)


def _seq(*mods):
    return _Tree(*[(i, m) for i, m in enumerate(mods) if m is not None])


def _down2(cin, cout):      # ResidualBlockWithStride2, layers.py:81-85
    return _Tree(("down", _c(cin, cout, 2, stride=2)), ("conv", _dcb(cout, cout)))


def _up2(cin, cout):        # ResidualBlockUpsample + SubpelConv2x, layers.py:22-29,93-97
    return _Tree(("up", _Tree(("conv", _seq(_c(cin, cout * 4, 1))))), ("conv", _dcb(cout, cout)))


class _Bitparm(nn.Module):
    """Parameters of Bitparm, entropy_models.py:84-97 (normal(0, 0.01) init; the final one has no `a`)."""

    def __init__(self, qp_num, ch, final):
        super().__init__()
        def table():
            return nn.Parameter(torch.nn.init.normal_(torch.empty(qp_num, ch, 1, 1), 0, 0.01))
        self.h = table()
        self.b = table()
        if not final:
            self.a = table()


def _bit_estimator(qp_num, ch):
    """BitEstimator, entropy_models.py:129-137."""
    return _Tree(("f1", _Bitparm(qp_num, ch, False)), ("f2", _Bitparm(qp_num, ch, False)),
                 ("f3", _Bitparm(qp_num, ch, False)), ("f4", _Bitparm(qp_num, ch, True)))


# --------------------------------------------------------------------------
# engine plumbing shared by all models
# --------------------------------------------------------------------------
# Any nn.Module.register_parameter anywhere (also `module.weight = nn.Parameter(...)`) bumps this counter; the
# modules below rebuild their cached parameter list when it moved, so a re-registered Parameter is never served from
# a stale list.
_param_generation = [0]


def _on_register_parameter(module, name, param):
    _param_generation[0] += 1
    return None


torch.nn.modules.module.register_module_parameter_registration_hook(_on_register_parameter)


class NonFiniteError(RuntimeError):
    """The reference's `[NaNGuard]` RuntimeError (seg_video_model_fast.py:152-156)."""


class _EngineModule(nn.Module):
    variant = "old"
    #: engine flags (see include/dmc_b200.h); tests switch these per instance
    engine_flags = 0
    #: engines kept alive per module, one per (B, H, W, device, flags); the least recently used one is destroyed
    #: beyond this (each holds 1.6-4 GB of workspace at 1920x1280)
    max_engines = 4
    #: every N-th forward the parameters are also checksummed on the device.  Writes through `.data`
    #: (`p.data.copy_(...)`, trainer_seg_video_model.py:789-791) change neither the storage pointer nor the version
    #: counter; they are caught here, or at once by calling `invalidate_weights()` after such a write.
    checksum_every = 32
    #: True: read the finite flag back after every forward and raise like the reference's _finite_check (one host
    #: sync per frame).  False (default): the flag is copied to pinned memory asynchronously and examined without
    #: blocking on later calls / by `check_finite()`, so the error surfaces at most a few frames late.
    strict_finite = os.environ.get("DMC_B200_STRICT_FINITE", "0") == "1"

    def _init_engine_state(self):
        self.__dict__["_engines"] = collections.OrderedDict()     # key -> handle, least recently used first
        self.__dict__["_weights_sig"] = {}
        self.__dict__["_weights_sum"] = {}
        self.__dict__["_lib"] = None
        self.__dict__["_param_cache"] = None
        self.__dict__["_param_gen"] = -1
        self.__dict__["_calls"] = 0
        self.__dict__["_finite"] = None

    # -- lifetime ----------------------------------------------------------
    def __del__(self):
        try:
            self.release_engines()
        except Exception:
            pass

    def release_engines(self):
        """Destroys every engine of this module (their workspaces are freed); the next forward rebuilds what it needs.
        Call it when the working resolution changes for good."""
        lib = self.__dict__.get("_lib")
        for h in self.__dict__.get("_engines", {}).values():
            if lib is not None:
                lib.dmc_destroy(h)
        self._init_engine_state()

    def invalidate_weights(self):
        """Forces a repack of the weights on the next forward (use after writing parameters through `.data`)."""
        self.__dict__["_weights_sig"] = {}
        self.__dict__["_param_cache"] = None

    # engine handles, the ctypes library and the caches are process-local: copies / pickles start without them
    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ("_engines", "_weights_sig", "_weights_sum", "_lib", "_param_cache", "_param_gen", "_calls", "_finite"):
            state.pop(k, None)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._init_engine_state()

    # -- helpers -----------------------------------------------------------
    def _check_mode(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("dmc_b200: CUDA tensors required (this implementation has no CPU path)")
        if self.training or torch.is_grad_enabled():
            raise NotImplementedError(
                "dmc_b200 implements the inference forward only: call model.eval() and wrap the call in "
                "torch.no_grad(). Training (STE / noise quantisation, inference.py:16-27) stays on the "
                "reference modules.")
        if x.dtype != torch.float32:
            raise TypeError("dmc_b200: float32 input expected")

    @staticmethod
    def _check_tensor(name: str, t: torch.Tensor, shape, device):
        """Same-device float32 tensor of the exact shape, or the shape error the reference's first conv would raise:
        the kernels read raw pointers, a wrong tensor would be an out-of-bounds access instead of an exception."""
        if not torch.is_tensor(t):
            raise TypeError(f"dmc_b200: {name} must be a tensor, got {type(t).__name__}")
        if t.device != device:
            raise RuntimeError(f"dmc_b200: {name} is on {t.device}, the input is on {device}")
        if t.dtype != torch.float32:
            raise TypeError(f"dmc_b200: {name} must be float32, got {t.dtype}")
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"dmc_b200: {name} has shape {tuple(t.shape)}, expected {tuple(shape)}")

    def _params(self):
        cache = self.__dict__.get("_param_cache")
        if cache is None or self.__dict__.get("_param_gen") != _param_generation[0]:
            cache = list(self.parameters())
            self.__dict__["_param_cache"] = cache
            self.__dict__["_param_gen"] = _param_generation[0]
        return cache

    def _signature(self):
        """(storage pointer, version counter) of every parameter: in-place edits bump the version, `.data = ...`
        swaps change the pointer, re-registration rebuilds the list (global registration hook), `.to()` /
        `load_state_dict` drop it.  What none of these see -- writes THROUGH `.data` -- is covered by the periodic
        device checksum in `_engine` and by `invalidate_weights()`."""
        return tuple((p.data_ptr(), p._version) for p in self._params())

    def _checksum(self):
        ps = [p.detach() for p in self._params() if p.is_cuda]
        if not ps:
            return None
        norms = torch.stack(torch._foreach_norm(ps)).double()
        return float((norms * torch.arange(1, len(ps) + 1, device=norms.device, dtype=torch.float64)).sum())

    def _apply(self, fn, *args, **kwargs):
        self.__dict__["_param_cache"] = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate_weights()
        return super().load_state_dict(*args, **kwargs)

    def _engine(self, B: int, H: int, W: int, device: torch.device):
        if self.__dict__.get("_lib") is None:
            self.__dict__["_lib"] = _capi.load()
        lib = self._lib
        key = (B, H, W, device.index, int(self.engine_flags))
        engines = self._engines
        h = engines.get(key)
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        if h is None:
            while len(engines) >= max(1, int(self.max_engines)):
                old_key, old_h = engines.popitem(last=False)
                lib.dmc_destroy(old_h)                     # (cudaFree synchronises: nothing of it is in flight)
                self._weights_sig.pop(old_key, None)
                self._weights_sum.pop(old_key, None)
            out = ctypes.c_void_p()
            with torch.cuda.device(device):
                rc = lib.dmc_create(_capi.VARIANT_IDS[self.variant], B, H, W, int(self.engine_flags),
                                    ctypes.byref(out))
            _capi.check(rc, None)
            h = out.value
            engines[key] = h
        else:
            engines.move_to_end(key)
        self.__dict__["_calls"] += 1
        sig = self._signature()
        stale = self._weights_sig.get(key) != sig
        if not stale and self.checksum_every and self._calls % int(self.checksum_every) == 0:
            stale = self._checksum() != self._weights_sum.get(key)
        if stale:
            sd = self.state_dict()
            keep = []                      # converted copies must outlive the async repack kernels
            for i in range(lib.dmc_num_weights(h)):
                name = lib.dmc_weight_key(h, i).decode()
                t = sd[name].detach()
                if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.to(device=device, dtype=torch.float32).contiguous()
                    keep.append(t)
                shape = (ctypes.c_int64 * t.dim())(*t.shape)
                rc = lib.dmc_set_weight(h, name.encode(), ctypes.c_void_p(t.data_ptr()), shape, t.dim(),
                                        stream)
                _capi.check(rc, h)
            if keep:
                torch.cuda.current_stream(device).synchronize()
            _capi.check(lib.dmc_finalize_weights(h, stream), h)
            self._weights_sig[key] = sig
            self._weights_sum[key] = self._checksum() if self.checksum_every else None
        return h, stream

    # -- finite flag (the reference's _finite_check, seg_video_model_fast.py:152-156) ------------------
    class _FiniteRing:
        """Device flag words + pinned mirrors + events: the check of forward n is examined, without blocking, during
        forward n+1, n+2, ... (or by check_finite()), so it costs no host sync."""
        SLOTS = 4

        def __init__(self, device):
            self.dev = torch.zeros(self.SLOTS, dtype=torch.int32, device=device)
            self.host = torch.zeros(self.SLOTS, dtype=torch.int32).pin_memory()
            self.events = [None] * self.SLOTS
            self.calls = [0] * self.SLOTS
            self.next = 0

    def _finite_slot(self, device):
        ring = self.__dict__.get("_finite")
        if ring is None or ring.dev.device != device:
            ring = self._FiniteRing(device)
            self.__dict__["_finite"] = ring
        self._poll_finite(block_slot=ring.next)            # the slot about to be reused must have been examined
        return ring, ring.next

    def _finite_submit(self, ring, slot):
        ring.host[slot:slot + 1].copy_(ring.dev[slot:slot + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(ring.dev.device))
        ring.events[slot] = ev
        ring.calls[slot] = self._calls
        ring.next = (slot + 1) % ring.SLOTS
        if self.strict_finite:
            self._poll_finite(block_slot=slot)

    def _poll_finite(self, block_slot=None, block_all=False):
        ring = self.__dict__.get("_finite")
        if ring is None:
            return
        for i, ev in enumerate(ring.events):
            if ev is None:
                continue
            if block_all or i == block_slot:
                ev.synchronize()
            elif not ev.query():
                continue
            ring.events[i] = None
            bits = int(ring.host[i])
            if bits:
                tags = ", ".join(t for b, t in enumerate(_capi.FINITE_TAGS) if bits >> b & 1)
                raise NonFiniteError(f"[NaNGuard] non-finite activations after {tags} (forward call {ring.calls[i]} of this "
                                     f"module; values at or beyond the fp16 range limit 65504 of the split storage "
                                     f"format count as non-finite)")

    def check_finite(self):
        """Blocks until every outstanding finite check of this module has been examined; raises like the reference's
        `_finite_check` if one of them failed."""
        self._poll_finite(block_all=True)

    def get_tap(self, name: str, x_like: torch.Tensor) -> torch.Tensor:
        """Intermediate tensor of the last forward (engine_flags must include FLAG_KEEP_TAPS)."""
        B, _, H, W = x_like.shape
        h, stream = self._engine(B, H, W, x_like.device)
        shape = (ctypes.c_int64 * 4)()
        with torch.cuda.device(x_like.device):
            _capi.check(self._lib.dmc_get_tap(h, name.encode(), None, 0, shape, stream), h)
            out = torch.empty(tuple(shape), dtype=torch.float32, device=x_like.device)
            _capi.check(self._lib.dmc_get_tap(h, name.encode(), ctypes.c_void_p(out.data_ptr()), out.numel(),
                                              shape, stream), h)
        return out


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p()


class _DMCBase(_EngineModule):
    """Common part of the four P-frame models."""

    def __init__(self, cfg: DMCConfig, refactor: bool):
        super().__init__()
        self.cfg = cfg
        self.qp_shift = list(cfg.qp_shift) if not refactor else cfg.qp_shift
        qp_num = 64 + cfg.extra_qp
        d, y, z, r = cfg.ch_d, cfg.ch_y, cfg.ch_z, cfg.ch_recon
        # CompressionModel.__init__ (common_model.py:15-23) registers the bit estimator first
        self.bit_estimator_z = _bit_estimator(qp_num, z)
        self.feature_adaptor_i = _dcb(cfg.src, d)
        self.feature_adaptor_p = _c(d, d)
        self.feature_extractor = _Tree(("conv1", _seq(_dcb(d, d), _dcb(d, d))),
                                       ("conv2", _seq(_dcb(d, d), _dcb(d, d), _dcb(d, d), _dcb(d, d))))
        if refactor:
            self.encoder = _Tree(("conv1", _c(cfg.src, d)),
                                 ("conv2", _seq(_dcb(2 * d, d), _dcb(d, d), _dcb(d, d))),
                                 ("down", _c(d, y, 3, stride=2, padding=1)))
        else:
            self.encoder = _Tree(("conv1", _c(cfg.src, d)), ("conv2", _seq(_dcb(2 * d, d), _dcb(d, d))),
                                 ("conv3", _dcb(d, d)), ("down", _c(d, y, 3, stride=2, padding=1)))
        self.hyper_encoder = _Tree(("conv", _seq(_dcb(y, z), _down2(z, z), _down2(z, z))))
        self.hyper_decoder = _Tree(("conv", _seq(_up2(z, z), _up2(z, z), _dcb(z, y))))
        self.temporal_prior_encoder = _down2(d, 2 * y)
        self.y_prior_fusion = _Tree(("conv", _seq(_dcb(3 * y, 3 * y), _dcb(3 * y, 3 * y),
                                                  _dcb(3 * y, 3 * y), _c(3 * y, 3 * y))))
        self.y_spatial_prior = _Tree(("conv", _seq(_dcb(4 * y, 3 * y), _dcb(3 * y, 3 * y),
                                                   _c(3 * y, 2 * y))))
        up = _Tree(("conv", _seq(_c(y, 4 * d, 3, padding=1))))
        if refactor:
            self.decoder = _Tree(("up", up), ("conv", _seq(_dcb(2 * d, d), _dcb(d, d), _dcb(d, d))),
                                 ("proj", _c(d, d)))
        else:
            self.decoder = _Tree(("up", up), ("conv1", _seq(_dcb(2 * d, d), _dcb(d, d), _dcb(d, d))),
                                 ("conv2", _c(d, d)))
        self.recon_generation_net = _Tree(("conv", _seq(_dcb(d, r), _dcb(r, r), _dcb(r, r), _dcb(r, r))),
                                          ("head", _c(r, cfg.src)))
        self._init_engine_state()

    def _register_q_tables(self):
        cfg = self.cfg
        qp_num = 64 + cfg.extra_qp
        self.q_encoder = nn.Parameter(torch.ones((qp_num, cfg.ch_d, 1, 1)))
        self.q_decoder = nn.Parameter(torch.ones((qp_num, cfg.ch_d, 1, 1)))
        self.q_feature = nn.Parameter(torch.ones((qp_num, cfg.ch_d, 1, 1)))
        self.q_recon = nn.Parameter(torch.ones((qp_num, cfg.ch_recon, 1, 1)))

    @staticmethod
    def get_qp_num():
        return 64

    def shift_qp(self, qp, fa_idx):
        """video_model.py:335-336."""
        return qp + self.qp_shift[fa_idx]

    def forward(self, x, qp, dpb, after_i=True):
        self._check_mode(x)
        B, C, H, W = x.shape
        takes_mask = self.variant != "old"
        if C == 3:
            x_img, mask = x, None
        elif C == 4 and takes_mask:
            x_img, mask = x[:, :3].contiguous(), x[:, 3:4].contiguous()
        else:
            # same failure as the reference's encoder.conv1 on a wrong channel count
            raise RuntimeError(f"expected input with 3{' or 4' if takes_mask else ''} channels, got {C}")
        # y is replicate-padded to multiples of 4 for the hyper path (old / fast / mask_prop); `performance` does not
        # pad (seg_video_model.py:331) and fails in the reference on such sizes too
        need = 64 if self.variant == "performance" else 16
        if H % need or W % need:
            raise RuntimeError(f"dmc_b200: height and width must be multiples of {need} for variant {self.variant}")
        x_img = x_img.contiguous()
        dev = x.device
        frame = feature = None
        if after_i:
            frame = dpb["frame"]
            self._check_tensor("dpb['frame']", frame, (B, 3, H, W), dev)
            frame = frame.contiguous()
        else:
            if dpb.get("feature") is None:
                raise RuntimeError("dpb['feature'] is None on a non-first P frame")
            feature = dpb["feature"]
            self._check_tensor("dpb['feature']", feature, (B, self.cfg.ch_d, H // 8, W // 8), dev)
            feature = feature.contiguous()
        h, stream = self._engine(B, H, W, dev)
        x_hat = torch.empty((B, 3, H, W), dtype=torch.float32, device=dev)
        feat = torch.empty((B, self.cfg.ch_d, H // 8, W // 8), dtype=torch.float32, device=dev)
        bpp3 = torch.empty((B, 3), dtype=torch.float32, device=dev)
        mask_pred = None
        if self.variant == "mask_prop" and not after_i and mask is not None:
            mask_pred = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
        ring, slot = self._finite_slot(dev)
        with torch.cuda.device(dev):
            rc = self._lib.dmc_forward(h, _ptr(x_img), _ptr(mask), _ptr(frame), _ptr(feature), int(qp),
                                       1 if after_i else 0, _ptr(x_hat), _ptr(feat), _ptr(bpp3),
                                       _ptr(mask_pred), ctypes.c_void_p(ring.dev.data_ptr() + 4 * slot), stream)
            _capi.check(rc, h)
            self._finite_submit(ring, slot)
        out = {"dpb": {"frame": x_hat, "feature": feat}, "bpp": bpp3[:, 0], "bpp_y": bpp3[:, 1],
               "bpp_z": bpp3[:, 2]}
        if self.variant == "fast":
            out["mask_pred"] = mask if not after_i else None
        elif self.variant == "mask_prop":
            out["mask_pred"] = mask_pred if not after_i else None
        return out


class DMC_old(_DMCBase):
    """`dmc_variant=old`: src/models/video_model.py:183-388 (constructor takes no arguments)."""
    variant = "old"

    def __init__(self):
        super().__init__(DMCConfig(), refactor=False)
        self._register_q_tables()


class DMC_performance(_DMCBase):
    """`dmc_variant=performance`: src/refactor/seg_video_model.py:205-365."""
    variant = "performance"

    def __init__(self, cfg: Optional[DMCConfig] = None):
        cfg = cfg or DMCConfig()
        super().__init__(cfg, refactor=True)
        self.hyper_in_adapter = _c(cfg.ch_y + 1, cfg.ch_y)          # unused by forward (:225)
        self.mask_sft = _Tree(("conv1", _c(cfg.patch_size ** 2, cfg.ch_d)),
                              ("conv2", _seq(_dcb(cfg.ch_d, cfg.ch_d), _dcb(cfg.ch_d, cfg.ch_d),
                                             _dcb(cfg.ch_d, cfg.ch_d))),
                              ("down", _c(cfg.ch_d, cfg.ch_y * 2, 3, stride=2, padding=1)))
        self._register_q_tables()
        self.q_sft = nn.Parameter(torch.ones((64 + cfg.extra_qp, cfg.ch_d, 1, 1)))


def _identity_adapter(cfg: DMCConfig):
    """seg_video_model_fast.py:205-206,255-265: identity on y, zero on the mask channel."""
    conv = _c(cfg.ch_y + 1, cfg.ch_y)
    with torch.no_grad():
        conv.weight.zero_()
        conv.bias.zero_()
        for i in range(cfg.ch_y):
            conv.weight[i, i, 0, 0] = 1.0
    return conv


def _mask_film(cfg: DMCConfig):
    return _Tree(("net", _Tree((0, _c(1, 16, 3, padding=1)), (2, _c(16, 2 * cfg.ch_y)))))


class DMC_fast(_DMCBase):
    """`dmc_variant=fast`: src/refactor/seg_video_model_fast.py:185-411."""
    variant = "fast"

    def __init__(self, cfg: Optional[DMCConfig] = None):
        cfg = cfg or DMCConfig()
        super().__init__(cfg, refactor=True)
        self.hyper_in_adapter = _identity_adapter(cfg)
        self.mask_film = _mask_film(cfg)
        self._register_q_tables()


class DMC_mask_prop(_DMCBase):
    """`dmc_variant=mask_prop`: src/refactor/mask_prop_seg_video_model.py:185-417."""
    variant = "mask_prop"

    def __init__(self, cfg: Optional[DMCConfig] = None):
        cfg = cfg or DMCConfig()
        super().__init__(cfg, refactor=True)
        self.hyper_in_adapter = _identity_adapter(cfg)
        self.mask_film = _mask_film(cfg)
        d = cfg.ch_d
        self.mask_predictor = _Tree(("mask_embed", _c(1, d, 3, padding=1)),          # mask_predictor.py:12-25
                                    ("net", _Tree((0, _c(3 * d, d // 4, 3, padding=1)),
                                                  (2, _c(d // 4, d // 4, 3, padding=1)),
                                                  (4, _c(d // 4, 1)))))
        self._register_q_tables()


class DMCI(_EngineModule):
    """Intra model: src/models/image_model.py:96-261."""
    variant = "intra"

    def __init__(self, N: int = 256, z_channel: int = 128):
        super().__init__()
        if (N, z_channel) != (256, 128):
            raise NotImplementedError("dmc_b200 DMCI is built for N=256, z_channel=128")
        e = 368
        self.bit_estimator_z = _bit_estimator(64, z_channel)
        self.enc = _Tree(("enc_1", _dcb(192, e)),
                         ("enc_2", _seq(*[_dcb(e, e) for _ in range(6)], _c(e, N, 3, stride=2, padding=1))))
        self.hyper_enc = _seq(_dcb(N, z_channel), _down2(z_channel, z_channel), _down2(z_channel, z_channel))
        self.hyper_dec = _seq(_up2(z_channel, z_channel), _up2(z_channel, z_channel), _dcb(z_channel, N))
        self.y_prior_fusion = _seq(_dcb(N, 2 * N), _dcb(2 * N, 2 * N), _dcb(2 * N, 2 * N),
                                   _c(2 * N, 2 * N + 2))
        self.y_spatial_prior_reduction = _c(2 * N + 2, N)
        self.y_spatial_prior_adaptor_1 = _dcb(2 * N, 2 * N, force_adaptor=True)
        self.y_spatial_prior_adaptor_2 = _dcb(2 * N, 2 * N, force_adaptor=True)
        self.y_spatial_prior_adaptor_3 = _dcb(2 * N, 2 * N, force_adaptor=True)
        self.y_spatial_prior = _seq(_dcb(2 * N, 2 * N), _dcb(2 * N, 2 * N), _dcb(2 * N, 2 * N),
                                    _c(2 * N, 2 * N))
        self.dec = _Tree(("dec_1", _seq(_up2(N, e), *[_dcb(e, e) for _ in range(12)])),
                         ("dec_2", _dcb(e, 192)))
        self.q_scale_enc = nn.Parameter(torch.ones((64, e, 1, 1)))
        self.q_scale_dec = nn.Parameter(torch.ones((64, e, 1, 1)))
        self._init_engine_state()

    @staticmethod
    def get_qp_num():
        return 64

    def forward(self, x, qp):
        self._check_mode(x)
        B, C, H, W = x.shape
        if C != 3:
            raise RuntimeError(f"expected a 3-channel frame, got {C}")
        if H % 16 or W % 16:
            raise RuntimeError("dmc_b200: height and width must be multiples of 16")
        x = x.contiguous()
        h, stream = self._engine(B, H, W, x.device)
        x_hat = torch.empty_like(x)
        bpp3 = torch.empty((B, 3), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _capi.check(self._lib.dmci_forward(h, _ptr(x), int(qp), _ptr(x_hat), _ptr(bpp3), stream), h)
        return {"dpb": {"frame": x_hat, "feature": None}, "bpp": bpp3[:, 0], "bpp_y": bpp3[:, 1],
                "bpp_z": bpp3[:, 2], "bits_y": torch.Size((B, 256, H // 16, W // 16)),
                "bits_z": torch.Size((B, 128, H // 64, W // 64))}


P_MODELS = {"old": DMC_old, "performance": DMC_performance, "fast": DMC_fast, "mask_prop": DMC_mask_prop}


def build_p_model(dmc_variant: str):
    """Mirror of the variant switch in trainer_seg_video_model.py:478-495."""
    if dmc_variant not in P_MODELS:
        raise ValueError(f"unknown dmc_variant {dmc_variant!r}")
    cls = P_MODELS[dmc_variant]
    return cls() if dmc_variant == "old" else cls(DMCConfig())
