"""Training-mode building blocks on the CUDA engine.

`p_frame_model.train()` in the reference (trainer_seg_video_model.py:983-1206) back-propagates the rate-distortion loss
through DepthConvBlock stacks, the STE / additive-noise quantisers and the Gaussian likelihood.  This module gives those
three pieces as `torch.autograd.Function`s over the C ABI (include/dmc_b200.h, "training mode"), wrapped in modules with
the reference's constructors and state_dict keys:

    DepthConvBlock(in_ch, out_ch, shortcut=False, force_adaptor=False)      src/layers/layers.py:43-79
    Conv2d(...)   nn.Conv2d subclass: dense 1x1 and k x k (2 / 3, stride 1 / 2) on the engine, the rest torch's
    AdaptiveQuant(mode="ste" | "noise", half_bin=0.5)                       src/layers/inference.py:8-27
    gaussian_bits(y, sigma, formula)                                        src/models/common_model.py:36-42 (0),
                                                                            src/refactor/common_model.py:37-68 (1)

Forward values are the inference engine's (same kernels); backward is hand-written CUDA: data gradients on the
tcgen05 chain kernel with transposed weights, weight gradients by a pixel-axis contraction, everything else
elementwise (csrc/train.cu).  Only x is kept between forward and backward -- the block's intermediates are recomputed.
There is no torch / CPU fallback: without the extension or a CUDA tensor the calls raise.
"""
from __future__ import annotations

import ctypes
import os
from collections import OrderedDict
from typing import Optional, Sequence

import torch
from torch import nn

from . import _capi

__all__ = ["DepthConvBlock", "AdaptiveQuant", "depth_conv_block", "gaussian_bits", "quant_ste", "quant_noise",
           "release_handles", "invalidate", "reference_patched", "adopt", "Conv2d"]

#: DepthConvBlock handles kept alive -- one per (owner module, geometry), each with its own workspace (~1.2 GB at
#: 160x240x256, 1/4 of that per halving of the resolution) and packed weights; least recently used first out
max_handles = 96

#: True (or DMC_B200_STRICT_FINITE=1): every block output is checked for non-finite / fp16-saturated values and the
#: reference's NaNGuard error is raised at the call (one host sync per block).  Activations pass through fp16 split planes:
#: |x| must stay below 65 504 (the conversions saturate).  Off by default, like the inference modules' lazy check.
strict_finite = os.environ.get("DMC_B200_STRICT_FINITE", "0") == "1"

_handles: "OrderedDict[tuple, int]" = OrderedDict()
_packed_sig: dict = {}        # handle key -> signature of the parameter values the handle has packed


def _signature(w12):
    """(storage pointer, version counter) per parameter: moves with every in-place update an optimizer makes.  Writes
    through `.data` bypass the counter (see modules.py); call `invalidate()` after such a write."""
    return tuple((w.data_ptr(), w._version) for w in w12 if w is not None)


def invalidate():
    """Forget which parameter values the handles have packed (after writes through `.data`)."""
    _packed_sig.clear()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dmc_b200.training: CUDA tensors required (there is no CPU path)")


def _dense(t: torch.Tensor) -> torch.Tensor:
    """fp32, contiguous, 16-byte aligned (a contiguous view at an odd storage offset is copied)."""
    t = t.contiguous().float()
    return t.clone() if t.data_ptr() % 16 else t


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr() if t is not None else None
    return arr


def _destroy(key, h):
    lib = _capi.load()
    fn = {"conv1x1": lib.dmc_conv1x1_train_destroy, "convkxk": lib.dmc_convkxk_train_destroy}.get(key[0],
                                                                                                  lib.dmc_dcb_train_destroy)
    fn(ctypes.c_void_p(h))


def release_handles():
    """Frees every cached block / convolution handle and workspace."""
    while _handles:
        key, h = _handles.popitem(last=False)
        _destroy(key, h)
    _packed_sig.clear()


def _handle(owner, device, B, H, W, cin, cout, force_adaptor, shortcut, has_qs, terms):
    """-> (key, handle).  `owner` separates blocks of equal geometry so that each keeps its own packed weights."""
    key = (owner, device.index, B, H, W, cin, cout, bool(force_adaptor), bool(shortcut), bool(has_qs), terms)
    lib = _capi.load()
    if key in _handles:
        _handles.move_to_end(key)
        return key, ctypes.c_void_p(_handles[key])
    while len(_handles) >= max_handles:
        old_key, old = _handles.popitem(last=False)
        _packed_sig.pop(old_key, None)
        _destroy(old_key, old)
    h = ctypes.c_void_p()
    with torch.cuda.device(device):
        rc = lib.dmc_dcb_train_create(B, H, W, cin, cout, int(force_adaptor), int(shortcut), int(has_qs), terms,
                                      ctypes.byref(h))
    if rc != 0:
        msg = lib.dmc_dcb_train_last_error(None)
        raise _capi.EngineError(f"dmc_dcb_train_create: {msg.decode() if msg else rc}")
    _handles[key] = h.value
    return key, h


def _check(rc, h):
    if rc != 0:
        msg = _capi.load().dmc_dcb_train_last_error(h)
        raise _capi.EngineError(f"dmc_b200 training error {rc}: {msg.decode() if msg else '?'}")


class _DepthConvBlockFn(torch.autograd.Function):
    """y = DepthConvBlock(x) [* quant_step];  inputs: x, quant_step or None, then the 12 parameters (adaptor.weight,
    adaptor.bias, dc.0, dc.2, dc.3, ffn.0, ffn.2; None for an absent adaptor)."""

    @staticmethod
    def forward(ctx, x, quant_step, shortcut, terms, owner, *w12):
        _need_cuda(x)
        lib = _capi.load()
        x = _dense(x)
        B, cin, H, W = x.shape
        cout = w12[2].shape[0]
        has_ad = w12[0] is not None
        ws = [None if w is None else w.detach().contiguous().float() for w in w12]
        qs = None
        if quant_step is not None:
            if quant_step.numel() != cout:
                raise RuntimeError("quant_step must hold one value per output channel")
            qs = quant_step.detach().reshape(cout).contiguous().float()
        key, h = _handle(owner, x.device, B, H, W, cin, cout, has_ad, shortcut, qs is not None, terms)
        sig = _signature(w12)
        unchanged = int(_packed_sig.get(key) == sig)
        out = torch.empty(B, cout, H, W, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _check(lib.dmc_dcb_train_forward(h, _ptr(x), _ptr_array(ws), _ptr(qs), _ptr(out), unchanged,
                                             _stream(x.device)), h)
        _packed_sig[key] = sig
        if strict_finite:
            peak = float(out.abs().max())
            if not peak < 65504.0:
                from .modules import NonFiniteError
                raise NonFiniteError(f"[NaNGuard] non-finite activations after DepthConvBlock({cin}->{cout}) "
                                     f"(max |x| = {peak}; the engine's fp16 split planes saturate at 65504)")
        # (the output is kept only when quant_step needs a gradient: d out / d quant_step = out / quant_step)
        keep_out = quant_step is not None and ctx.needs_input_grad[1]
        ctx.save_for_backward(x, quant_step, out if keep_out else None, *[w for w in w12 if w is not None])
        ctx.meta = (has_ad, bool(shortcut), terms, quant_step is not None, owner, sig)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        has_ad, shortcut, terms, has_qs, owner, sig = ctx.meta
        saved = ctx.saved_tensors
        x, quant_step, out = saved[0], saved[1], saved[2]
        w12 = list(saved[3:])
        if not has_ad:
            w12 = [None, None] + w12
        lib = _capi.load()
        B, cin, H, W = x.shape
        cout = w12[2].shape[0]
        ws = [None if w is None else w.detach().contiguous().float() for w in w12]
        qs = quant_step.detach().reshape(cout).contiguous().float() if has_qs else None
        g = _dense(grad_out)                    # (scaled into fp16's range inside the engine)
        need = ctx.needs_input_grad            # (x, quant_step, shortcut, terms, owner, *w12)
        gx = torch.empty_like(x) if need[0] else None
        sizes = [0 if (w is None or not need[5 + i]) else w.numel() for i, w in enumerate(w12)]
        n_qs = cout if (has_qs and need[1]) else 0
        flat = torch.empty(sum(sizes) + n_qs, device=x.device, dtype=torch.float32)
        gws, off = [], 0
        for n in sizes:
            gws.append(flat[off:off + n] if n else None)
            off += n
        gqs = flat[off:off + n_qs] if n_qs else None
        key, h = _handle(owner, x.device, B, H, W, cin, cout, has_ad, shortcut, has_qs, terms)
        unchanged = int(_packed_sig.get(key) == sig)      # (autograd itself refuses saved tensors modified in place)
        with torch.cuda.device(x.device):
            _check(lib.dmc_dcb_train_backward(h, _ptr(x), _ptr_array(ws), _ptr(qs), _ptr(out), _ptr(g), _ptr(gx),
                                              _ptr_array(gws), _ptr(gqs), unchanged, _stream(x.device)), h)
        _packed_sig[key] = sig
        grads_w = [None if gw is None else gw.view_as(w) for gw, w in zip(gws, w12)]
        g_qs = gqs.view_as(quant_step) if gqs is not None else None
        return (gx, g_qs, None, None, None, *grads_w)


def depth_conv_block(x, weights12, quant_step=None, shortcut=False, terms=3, owner=0):
    """Functional form; `weights12` in the order of include/dmc_b200.h (adaptor entries None when the block has none).
    `owner`: any hashable that tells blocks of equal geometry apart (each then keeps its own packed weights)."""
    return _DepthConvBlockFn.apply(x, quant_step, bool(shortcut), int(terms), owner, *weights12)


class _Seq(nn.Module):
    """nn.Sequential-like container that keeps the reference's sparse child indices (dc.0, dc.2, dc.3; ffn.0, ffn.2)."""

    def __init__(self, children):
        super().__init__()
        for name, m in children:
            self.add_module(str(name), m)

    def __getitem__(self, i):
        return self._modules[str(i)]


class DepthConvBlock(nn.Module):
    """src/layers/layers.py:43-79 with the same constructor, parameters and forward signature; runs on the engine in
    train and eval mode alike (`terms`: 3 = fp32-grade products, 1 = plain fp16 operands)."""

    def __init__(self, in_ch, out_ch, shortcut=False, force_adaptor=False, terms=3):
        super().__init__()
        self.adaptor = None
        if in_ch != out_ch or force_adaptor:
            self.adaptor = nn.Conv2d(in_ch, out_ch, 1)
        self.shortcut = shortcut
        self.terms = terms
        self.dc = _Seq([(0, nn.Conv2d(out_ch, out_ch, 1)), (2, nn.Conv2d(out_ch, out_ch, 3, padding=1, groups=out_ch)),
                        (3, nn.Conv2d(out_ch, out_ch, 1))])
        self.ffn = _Seq([(0, nn.Conv2d(out_ch, out_ch * 4, 1)), (2, nn.Conv2d(out_ch * 2, out_ch, 1))])

    def weights12(self):
        ad = self.adaptor
        convs = [self.dc[0], self.dc[2], self.dc[3], self.ffn[0], self.ffn[2]]
        out = [ad.weight if ad is not None else None, ad.bias if ad is not None else None]
        for c in convs:
            out += [c.weight, c.bias]
        return out

    def forward(self, x, quant_step=None, to_cat=None, cat_at_front=True):
        if quant_step is not None and quant_step.numel() != self.dc[0].out_channels:
            raise RuntimeError("DepthConvBlock: quant_step must be one (1, C, 1, 1) row (a single qp per call)")
        out = depth_conv_block(x, self.weights12(), quant_step, self.shortcut, self.terms, owner=id(self))
        if to_cat is not None:
            out = torch.cat((to_cat, out), dim=1) if cat_at_front else torch.cat((out, to_cat), dim=1)
        return out

    forward_torch = forward        # the reference's forward() delegates to a method of this name


# ------------------------------------------------------------------------------------------------ plain 1x1 convolution
def _conv_handle(owner, device, B, H, W, cin, cout, has_bias, terms, geom=(1, 1, 0)):
    """geom = (kernel, stride, padding); (1, 1, 0) is the plain 1x1 handle, anything else the im2col one."""
    kxk = geom != (1, 1, 0)
    key = ("convkxk" if kxk else "conv1x1", owner, device.index, B, H, W, cin, cout, bool(has_bias), terms, geom)
    lib = _capi.load()
    if key in _handles:
        _handles.move_to_end(key)
        return key, ctypes.c_void_p(_handles[key])
    while len(_handles) >= max_handles:
        old_key, old = _handles.popitem(last=False)
        _packed_sig.pop(old_key, None)
        _destroy(old_key, old)
    h = ctypes.c_void_p()
    with torch.cuda.device(device):
        if kxk:
            rc = lib.dmc_convkxk_train_create(B, H, W, cin, cout, geom[0], geom[1], geom[2], int(has_bias), terms,
                                              ctypes.byref(h))
        else:
            rc = lib.dmc_conv1x1_train_create(B, H, W, cin, cout, int(has_bias), terms, ctypes.byref(h))
    if rc != 0:
        msg = (lib.dmc_convkxk_train_last_error if kxk else lib.dmc_conv1x1_train_last_error)(None)
        raise _capi.EngineError(f"conv train handle: {msg.decode() if msg else rc}")
    _handles[key] = h.value
    return key, h


def _check_conv(rc, h, kxk=False):
    if rc != 0:
        lib = _capi.load()
        msg = (lib.dmc_convkxk_train_last_error if kxk else lib.dmc_conv1x1_train_last_error)(h)
        raise _capi.EngineError(f"dmc_b200 training error {rc}: {msg.decode() if msg else '?'}")


class _Conv1x1Fn(torch.autograd.Function):
    """A dense convolution on the engine: geom = (kernel, stride, padding); (1, 1, 0) is the plain 1x1 case."""

    @staticmethod
    def forward(ctx, x, weight, bias, terms, owner, geom=(1, 1, 0)):
        _need_cuda(x)
        lib = _capi.load()
        x = _dense(x)
        B, cin, H, W = x.shape
        cout = weight.shape[0]
        k, stride, pad = geom
        kxk = geom != (1, 1, 0)
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        w = _dense(weight.detach())
        b = _dense(bias.detach()) if bias is not None else None
        key, h = _conv_handle(owner, x.device, B, H, W, cin, cout, b is not None, terms, geom)
        sig = _signature((weight, bias))
        unchanged = int(_packed_sig.get(key) == sig)
        out = torch.empty(B, cout, Ho, Wo, device=x.device, dtype=torch.float32)
        fwd = lib.dmc_convkxk_train_forward if kxk else lib.dmc_conv1x1_train_forward
        with torch.cuda.device(x.device):
            _check_conv(fwd(h, _ptr(x), _ptr(w), _ptr(b), _ptr(out), unchanged, _stream(x.device)), h, kxk)
        _packed_sig[key] = sig
        ctx.save_for_backward(x, weight)
        ctx.meta = (bias is not None, terms, owner, sig, geom)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        has_bias, terms, owner, sig, geom = ctx.meta
        x, weight = ctx.saved_tensors
        lib = _capi.load()
        B, cin, H, W = x.shape
        cout = weight.shape[0]
        kxk = geom != (1, 1, 0)
        g = _dense(grad_out)
        need = ctx.needs_input_grad            # (x, weight, bias, terms, owner, geom)
        gx = torch.empty_like(x) if need[0] else None
        nw = weight.numel() if need[1] else 0
        nb = cout if (has_bias and need[2]) else 0
        flat = torch.empty(nw + nb, device=x.device, dtype=torch.float32)
        gw = flat[:nw] if nw else None
        gb = flat[nw:] if nb else None
        key, h = _conv_handle(owner, x.device, B, H, W, cin, cout, has_bias, terms, geom)
        unchanged = int(_packed_sig.get(key) == sig)
        bwd = lib.dmc_convkxk_train_backward if kxk else lib.dmc_conv1x1_train_backward
        with torch.cuda.device(x.device):
            _check_conv(bwd(h, _ptr(x), _ptr(_dense(weight.detach())), _ptr(g), _ptr(gx), _ptr(gw), _ptr(gb), unchanged,
                            _stream(x.device)), h, kxk)
        _packed_sig[key] = sig
        return gx, (gw.view_as(weight) if gw is not None else None), gb, None, None, None


class Conv2d(nn.Conv2d):
    """nn.Conv2d whose dense instances the engine covers run on it, forward and backward (fp32-grade split products):
    1x1 / stride 1 / unpadded, and k x k with kernel 2 or 3, stride 1 or 2, padding 0 or 1 (through the im2col view) --
    ungrouped, undilated, zero padding, channel counts in multiples of 16 (cout >= 32; 1x1: cin >= 32).  Every other
    configuration is torch's own convolution, unchanged.  Same constructor, same parameters."""

    terms = 3
    #: False keeps the k x k instances on torch (A/B runs)
    kxk_on_engine = True

    def _geom(self, x):
        """(kernel, stride, padding) when this call runs on the engine, else None."""
        if not (x.is_cuda and x.dim() == 4 and self.dilation == (1, 1) and self.groups == 1
                and self.padding_mode == "zeros" and self.in_channels % 16 == 0 and self.out_channels % 16 == 0
                and self.out_channels >= 32 and isinstance(self.padding, tuple)):
            return None
        k, s, p = self.kernel_size, self.stride, self.padding
        if k[0] != k[1] or s[0] != s[1] or p[0] != p[1]:
            return None
        geom = (k[0], s[0], p[0])
        if geom == (1, 1, 0):
            return geom if self.in_channels >= 32 else None
        if self.kxk_on_engine and k[0] in (2, 3) and s[0] in (1, 2) and p[0] in (0, 1):
            return geom
        return None

    def _on_engine(self, x):
        return self._geom(x) is not None

    def forward(self, x):
        geom = self._geom(x)
        if geom is not None:
            return _Conv1x1Fn.apply(x, self.weight, self.bias, self.terms, id(self), geom)
        return super().forward(x)


class _NNProxy:
    """`nn` as the reference modules see it inside `reference_patched`: torch.nn with Conv2d replaced."""

    Conv2d = Conv2d

    def __getattr__(self, name):
        return getattr(nn, name)


# ------------------------------------------------------------------------------------------------ quantisation
class _QuantFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, noise, mode):
        _need_cuda(x, noise)
        lib = _capi.load()
        x = x.contiguous().float()
        out = torch.empty_like(x)
        nz = noise.contiguous().float() if noise is not None else None
        with torch.cuda.device(x.device):
            rc = lib.dmc_op_quant_train(_ptr(x), _ptr(nz), _ptr(out), x.numel(), mode, _stream(x.device))
        if rc != 0:
            raise _capi.EngineError(f"dmc_op_quant_train failed ({rc})")
        return out

    @staticmethod
    def backward(ctx, g):
        # ste: (round(x) - x).detach() + x  and  noise: x + noise  both have d out / d x = 1  (inference.py:18,25)
        return g, None, None


def quant_ste(x):
    return _QuantFn.apply(x, None, 0)


def quant_noise(x, half_bin=0.5, generator=None):
    noise = torch.empty_like(x).uniform_(-half_bin, half_bin, generator=generator)
    return _QuantFn.apply(x, noise, 1)


class AdaptiveQuant(nn.Module):
    """src/layers/inference.py:8-27.  Training: "ste" rounds with a straight-through gradient, "noise" adds
    U(-half_bin, half_bin) drawn from torch's generator in the reference's call order (a seeded run reproduces the
    reference's noise).  Eval: hard rounding in both modes."""

    def __init__(self, mode="ste", half_bin=0.5):
        super().__init__()
        assert mode in ["ste", "noise"], "Unsupported mode"
        self.mode = mode
        self.half_bin = half_bin

    def forward(self, x):
        if self.mode == "noise" and self.training:
            return quant_noise(x, self.half_bin)
        return quant_ste(x)


# ------------------------------------------------------------------------------------------------ likelihood
class _GaussianBitsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, sigma, formula):
        _need_cuda(y, sigma)
        lib = _capi.load()
        y = y.contiguous().float()
        sigma = sigma.contiguous().float()
        bits = torch.empty_like(y)
        with torch.cuda.device(y.device):
            rc = lib.dmc_op_gaussian_bits(_ptr(y), _ptr(sigma), _ptr(bits), y.numel(), formula, _stream(y.device))
        if rc != 0:
            raise _capi.EngineError(f"dmc_op_gaussian_bits failed ({rc})")
        ctx.save_for_backward(y, sigma)
        ctx.formula = formula
        return bits

    @staticmethod
    def backward(ctx, g):
        y, sigma = ctx.saved_tensors
        lib = _capi.load()
        g = g.contiguous().float()
        gy, gs = torch.empty_like(y), torch.empty_like(sigma)
        with torch.cuda.device(y.device):
            rc = lib.dmc_op_gaussian_bits_backward(_ptr(y), _ptr(sigma), _ptr(g), _ptr(gy), _ptr(gs), y.numel(),
                                                   ctx.formula, _stream(y.device))
        if rc != 0:
            raise _capi.EngineError(f"dmc_op_gaussian_bits_backward failed ({rc})")
        return gy, gs, None


def gaussian_bits(y, sigma, formula=1):
    """Per-element likelihood bits with gradients for y and sigma.  formula 0: models/common_model.py:36-42 (`old`,
    DMCI); 1: refactor/common_model.py:37-68 including the +-6 clamp of seg_video_model.py:347."""
    return _GaussianBitsFn.apply(y, sigma, int(formula))


# ------------------------------------------------------------------------------------------------ reference integration
class reference_patched:
    """Context manager: inside it, the given (already imported) reference modules construct the engine's blocks.

        import src.layers.layers as L, src.refactor.common_model as CM, src.refactor.seg_video_model as SV
        with dmc_b200.training.reference_patched(L, CM, SV):
            p_frame_model = SV.DMC(DMCConfig())          # DepthConvBlock / AdaptiveQuant / nn.Conv2d are now the engine's
        dmc_b200.training.adopt(p_frame_model, formula=1)   # + the likelihood with its native backward
        p_frame_model.load_state_dict(checkpoint)        # parameter names and shapes are the reference's

    Every name bound by `from ..layers.layers import DepthConvBlock` is a separate module attribute, so each module that
    constructs blocks has to be listed (layers.py itself for ResidualBlockWithStride2 / ResidualBlockUpsample)."""

    def __init__(self, *modules, convs=True):
        self.modules = modules
        self.convs = convs              # also build the models' nn.Conv2d layers as `Conv2d` (dense 1x1 / k x k on the engine)
        self.saved = []

    def __enter__(self):
        swaps = [("DepthConvBlock", DepthConvBlock), ("AdaptiveQuant", AdaptiveQuant)]
        for m in self.modules:
            for name, repl in swaps:
                if hasattr(m, name):
                    self.saved.append((m, name, getattr(m, name)))
                    setattr(m, name, repl)
            if self.convs and getattr(m, "nn", None) is nn:
                self.saved.append((m, "nn", nn))
                setattr(m, "nn", _NNProxy())
        return self

    def __exit__(self, *exc):
        for m, name, old in reversed(self.saved):
            setattr(m, name, old)
        self.saved = []
        return False


def adopt(model, formula=1):
    """Routes `model.get_y_gaussian_bits` (models/common_model.py:36-42 -> formula 0, refactor/common_model.py:37-68 ->
    formula 1) of this instance through the engine's likelihood kernels, forward and backward."""
    import types

    def get_y_gaussian_bits(self, y, sigma):
        return gaussian_bits(y, sigma, formula)

    model.get_y_gaussian_bits = types.MethodType(get_y_gaussian_bits, model)
    return model
