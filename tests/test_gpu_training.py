"""Training-mode pieces against torch.autograd of the reference's layer code, on a B200.

The checker is a plain torch restatement of src/layers/layers.py:43-79, src/layers/inference.py:16-27 and the two
`get_y_gaussian_bits` formulas, evaluated in fp64 on the CPU.  Tolerances (relative to the largest magnitude of the
tensor compared): forward 2e-5 (as the per-layer parity tests), gradients 1e-5 with the fp32-grade 3-term product
(measured on a B200: 1.3e-7 ... 6.5e-7 over all cases and tensors), 3e-3 with plain fp16 operands (`terms=1`; measured up to
6.2e-4).  Each is written next to its check.
"""
import math

import pytest
import torch
import torch.nn.functional as F
from torch import nn

from helpers import D

pytestmark = pytest.mark.gpu
T = D.training


def wsilu(x):
    return F.silu(4.0 * x) / 4.0


class RefDCB(nn.Module):
    """layers.py:43-79, verbatim semantics."""

    def __init__(self, cin, cout, shortcut=False, force_adaptor=False):
        super().__init__()
        self.adaptor = nn.Conv2d(cin, cout, 1) if (cin != cout or force_adaptor) else None
        self.shortcut = shortcut
        self.dc0, self.dc2, self.dc3 = nn.Conv2d(cout, cout, 1), nn.Conv2d(cout, cout, 3, padding=1, groups=cout), nn.Conv2d(cout, cout, 1)
        self.ffn0, self.ffn2 = nn.Conv2d(cout, cout * 4, 1), nn.Conv2d(cout * 2, cout, 1)

    def forward(self, x, quant_step=None):
        if self.adaptor is not None:
            x = self.adaptor(x)
        out = self.dc3(self.dc2(wsilu(self.dc0(x)))) + x
        u = wsilu(self.ffn0(out))
        u1, u2 = torch.chunk(u, 2, dim=1)
        out = self.ffn2(u1 + u2) + out
        if self.shortcut:
            out = out + x
        if quant_step is not None:
            out = out * quant_step
        return out


def _copy_params(ref: RefDCB, blk):
    pairs = [(ref.dc0, blk.dc[0]), (ref.dc2, blk.dc[2]), (ref.dc3, blk.dc[3]), (ref.ffn0, blk.ffn[0]), (ref.ffn2, blk.ffn[2])]
    if ref.adaptor is not None:
        pairs.append((ref.adaptor, blk.adaptor))
    with torch.no_grad():
        for r, b in pairs:
            b.weight.copy_(r.weight.float())
            b.bias.copy_(r.bias.float())
    return pairs


def _relmax(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


DCB_CASES = [
    # B, H, W, cin, cout, shortcut, force_adaptor, quant_step, grad magnitude
    (2, 16, 24, 64, 64, False, False, False, 1.0),
    (1, 20, 12, 48, 64, True, False, True, 1e-7),       # adaptor + shortcut + quant_step, tiny gradients (mean-reduced loss)
    (1, 16, 24, 128, 128, False, True, True, 3e3),      # forced adaptor, large gradients
    (1, 40, 60, 256, 256, False, False, True, 1e-5),    # the frame's block width (M = 2400: ragged 256-row tile)
    (1, 9, 7, 320, 320, True, False, False, 1.0),       # recon width: N = 2.5 weight-gradient tiles, M = 63
]


@pytest.mark.parametrize("case", DCB_CASES, ids=lambda c: "x".join(str(v) for v in c[:5]))
@pytest.mark.parametrize("terms", [3, 1])
def test_depth_conv_block_forward_backward(case, terms):
    B, H, W, cin, cout, shortcut, force, with_qs, gmag = case
    torch.manual_seed(1234 + cin + cout)
    ref = RefDCB(cin, cout, shortcut, force).double()
    blk = T.DepthConvBlock(cin, cout, shortcut, force, terms=terms).cuda()
    pairs = _copy_params(ref, blk)
    x = torch.randn(B, cin, H, W, dtype=torch.float64)
    qs = (torch.rand(1, cout, 1, 1, dtype=torch.float64) + 0.5) if with_qs else None
    gout = torch.randn(B, cout, H, W, dtype=torch.float64) * gmag

    xr = x.clone().requires_grad_(True)
    qr = qs.clone().requires_grad_(True) if with_qs else None
    yr = ref(xr, qr)
    yr.backward(gout)

    xg = x.float().cuda().requires_grad_(True)
    qg = qs.float().cuda().requires_grad_(True) if with_qs else None
    blk.train()
    yg = blk(xg, qg)
    yg.backward(gout.float().cuda())
    torch.cuda.synchronize()

    ftol = 2e-5 if terms == 3 else 2e-3
    gtol = 1e-5 if terms == 3 else 3e-3
    errs = {"out": (_relmax(yg, yr), ftol), "grad_x": (_relmax(xg.grad, xr.grad), gtol)}
    if with_qs:
        errs["grad_quant_step"] = (_relmax(qg.grad, qr.grad), gtol)
    names = ["dc.0", "dc.2", "dc.3", "ffn.0", "ffn.2", "adaptor"]
    for (r, b), n in zip(pairs, names):
        errs[f"grad_{n}.weight"] = (_relmax(b.weight.grad, r.weight.grad), gtol)
        errs[f"grad_{n}.bias"] = (_relmax(b.bias.grad, r.bias.grad), gtol)
    report = ", ".join(f"{k}={v[0]:.2e}" for k, v in errs.items())
    print(f"\nDCB {case} terms={terms}: {report}")
    bad = {k: v for k, v in errs.items() if not v[0] <= v[1]}
    assert not bad, f"outside tolerance: {bad}   (all: {report})"


def test_depth_conv_block_eval_matches_train_forward_and_keeps_state_dict_keys():
    blk = T.DepthConvBlock(48, 64).cuda()
    assert sorted(blk.state_dict()) == sorted(
        [f"{p}.{w}" for p in ("adaptor", "dc.0", "dc.2", "dc.3", "ffn.0", "ffn.2") for w in ("weight", "bias")])
    x = torch.randn(1, 48, 16, 16, device="cuda")
    blk.train()
    a = blk(x)
    blk.eval()
    with torch.no_grad():
        b = blk(x)
    assert torch.equal(a, b)           # the STE forward is the eval forward value for value
    cat = blk(x, to_cat=x, cat_at_front=False)
    assert cat.shape == (1, 64 + 48, 16, 16) and torch.equal(cat[:, :64], a)


def test_adaptive_quant_train_modes():
    x = (torch.randn(2, 8, 5, 7, device="cuda") * 3).requires_grad_(True)
    ste, noise = T.AdaptiveQuant("ste").cuda().train(), T.AdaptiveQuant("noise").cuda().train()
    y = ste(x)
    assert torch.equal(y, torch.round(x))                # inference.py:18
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))       # straight through
    x.grad = None
    torch.manual_seed(7)
    z = noise(x)
    torch.manual_seed(7)
    ref = x + torch.empty_like(x).uniform_(-0.5, 0.5)    # inference.py:23-25, same generator state
    assert torch.equal(z, ref)
    (z * 2).sum().backward()
    assert torch.equal(x.grad, torch.full_like(x, 2.0))
    noise.eval()
    assert torch.equal(noise(x), torch.round(x))         # eval: hard rounding in both modes


def _bits_ref(y, sigma, formula):
    if formula == 0:      # models/common_model.py:30-42
        sigma = sigma.clamp(1e-5, 1e10)
        g = torch.distributions.normal.Normal(torch.zeros_like(sigma), sigma)
        p = g.cdf(y + 0.5) - g.cdf(y - 0.5)
        return torch.clamp(torch.log(p + 1e-5) * (-1.0 / math.log(2.0)), 0, None)
    y = y.clamp(-6.0, 6.0)    # seg_video_model.py:347 + refactor/common_model.py:37-68
    sigma = sigma.clamp(1e-5, 1e10)
    inv = 1.0 / sigma
    zh, zl = ((y + 0.5) * inv).clamp(-12, 12), ((y - 0.5) * inv).clamp(-12, 12)
    p = 0.5 * (torch.erf(zh / math.sqrt(2.0)) - torch.erf(zl / math.sqrt(2.0)))
    return -torch.log2(p.clamp_min(1e-9))


@pytest.mark.parametrize("formula", [0, 1])
def test_gaussian_bits_backward(formula):
    torch.manual_seed(5 + formula)
    n = 1 << 16
    y = torch.round(torch.randn(n, dtype=torch.float64) * 2.5)
    y[::7] += torch.rand(n, dtype=torch.float64)[::7] - 0.5          # the noise quantiser's non-integers
    y[:64] = torch.linspace(-9, 9, 64, dtype=torch.float64)          # beyond the +-6 clamp
    sigma = torch.exp(torch.randn(n, dtype=torch.float64) * 1.2 - 0.3)
    sigma[64:96] = 1e-6                                              # below the sigma clamp
    go = torch.randn(n, dtype=torch.float64)
    yr, sr = y.clone().requires_grad_(True), sigma.clone().requires_grad_(True)
    _bits_ref(yr, sr, formula).backward(go)
    yg, sg = y.float().cuda().requires_grad_(True), sigma.float().cuda().requires_grad_(True)
    bits = T.gaussian_bits(yg, sg, formula)
    bits.backward(go.float().cuda())
    # where the reference's own fp32 forward saturates (p == 0 after cancellation) the gradient is not defined by the
    # formula but by rounding; compare where the fp64 likelihood is comfortably inside fp32's resolution
    with torch.no_grad():
        b64 = _bits_ref(y, sigma, formula)
        ok = b64 < 20.0
    gy, gs = yg.grad.double().cpu(), sg.grad.double().cpu()
    ey = ((gy - yr.grad).abs() / (yr.grad.abs() + 1e-3))[ok].max().item()
    es = ((gs - sr.grad).abs() / (sr.grad.abs() + 1e-3))[ok].max().item()
    print(f"\nbits backward formula {formula}: d/dy {ey:.2e}  d/dsigma {es:.2e}  ({int(ok.sum())} of {n} compared)")
    assert ey < 1e-4 and es < 1e-4      # fp32 inputs / outputs around an fp64 evaluation


def test_training_blocks_refuse_cpu_tensors():
    blk = T.DepthConvBlock(32, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        blk(torch.randn(1, 32, 8, 8))


def test_weight_updates_are_picked_up():
    """An optimizer step (in-place, bumps the version counter) must reach the packed copies; a write through `.data`
    does after `training.invalidate()`."""
    torch.manual_seed(3)
    blk = T.DepthConvBlock(64, 64).cuda().train()
    ref = RefDCB(64, 64).double()
    x = torch.randn(1, 64, 16, 16)

    def check():
        _copy_back(ref, blk)
        xr = x.double().requires_grad_(True)
        ref.zero_grad()
        ref(xr).sum().backward()
        xg = x.cuda().requires_grad_(True)
        blk.zero_grad()
        y = blk(xg)
        y.sum().backward()
        assert _relmax(y, ref(xr)) < 2e-5
        assert _relmax(blk.ffn[0].weight.grad, ref.ffn0.weight.grad) < 2e-4
        assert _relmax(xg.grad, xr.grad) < 2e-4

    def _copy_back(r, b):      # reference <- engine module (the engine module is the one being updated)
        with torch.no_grad():
            for rc, bc in [(r.dc0, b.dc[0]), (r.dc2, b.dc[2]), (r.dc3, b.dc[3]), (r.ffn0, b.ffn[0]), (r.ffn2, b.ffn[2])]:
                rc.weight.copy_(bc.weight.double().cpu())
                rc.bias.copy_(bc.bias.double().cpu())

    check()
    opt = torch.optim.SGD(blk.parameters(), lr=2e-5)   # (a sum loss: gradients are O(1e3))
    opt.step()                              # uses the gradients of check(): every parameter moves
    check()
    blk.ffn[0].weight.data.mul_(1.5)        # invisible to the version counter
    T.invalidate()
    check()


def test_strict_finite_raises_on_saturation(monkeypatch):
    monkeypatch.setattr(T, "strict_finite", True)
    blk = T.DepthConvBlock(32, 32).cuda()
    x = torch.randn(1, 32, 8, 8, device="cuda")
    blk(x)                                            # fine
    with torch.no_grad():
        blk.ffn[2].bias.fill_(1e6)                    # pushes the output beyond fp16's range
    with pytest.raises(D.NonFiniteError, match="NaNGuard"):
        blk(x)


def test_blocks_of_equal_geometry_share_a_workspace_but_not_weights():
    """Two blocks of the same geometry interleaved (forward A, forward B, backward B, backward A): the shared workspace
    must not leak one block's intermediates or packed weights into the other's results."""
    torch.manual_seed(9)
    a, b = T.DepthConvBlock(64, 64).cuda().train(), T.DepthConvBlock(64, 64).cuda().train()
    ra, rb = RefDCB(64, 64).double(), RefDCB(64, 64).double()
    for ref, blk in ((ra, a), (rb, b)):
        _copy_params(ref, blk)
    x = torch.randn(1, 64, 12, 20)
    xa, xb = x.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    ya = a(xa)
    yb = b(xb)
    (yb * 2).sum().backward()
    ya.sum().backward()
    for ref, blk, xg, k in ((ra, a, xa, 1.0), (rb, b, xb, 2.0)):
        xr = x.double().requires_grad_(True)
        (ref(xr) * k).sum().backward()
        assert _relmax(xg.grad, xr.grad) < 2e-4
        assert _relmax(blk.ffn[0].weight.grad, ref.ffn0.weight.grad) < 2e-4
        assert _relmax(blk.dc[2].weight.grad, ref.dc2.weight.grad) < 2e-4


def test_unaligned_and_strided_inputs_are_accepted():
    """A contiguous view at an odd storage offset (not 16-byte aligned) and a channels-last tensor."""
    torch.manual_seed(4)
    blk = T.DepthConvBlock(32, 32).cuda().train()
    base = torch.randn(1 + 32 * 6 * 7, device="cuda")
    x_odd = base[1:].view(1, 32, 6, 7)                      # data_ptr % 16 == 4
    assert x_odd.data_ptr() % 16 != 0 and x_odd.is_contiguous()
    x_ref = x_odd.clone().requires_grad_(True)
    x_cl = x_odd.clone().to(memory_format=torch.channels_last).requires_grad_(True)
    y_ref = blk(x_ref)
    g = torch.randn_like(y_ref)
    y_ref.backward(g)
    for xv in (x_odd.detach().requires_grad_(True), x_cl):
        blk.zero_grad()
        y = blk(xv)
        y.backward(g)
        assert torch.equal(y, y_ref)
        assert torch.allclose(xv.grad, x_ref.grad, rtol=0, atol=0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_block_on_a_device_that_is_not_current():
    """model.to('cuda:1') while cuda:0 stays the current device (what DDP ranks with CUDA_VISIBLE_DEVICES unset do)."""
    torch.manual_seed(6)
    ref = RefDCB(64, 64).double()
    x = torch.randn(1, 64, 10, 14, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    ref(xr).sum().backward()
    assert torch.cuda.current_device() == 0
    for dev in ("cuda:1", "cuda:0"):
        blk = T.DepthConvBlock(64, 64).to(dev).train()
        _copy_params(ref, blk)
        xg = x.float().to(dev).requires_grad_(True)
        y = blk(xg)
        y.sum().backward()
        assert y.device == torch.device(dev) and torch.cuda.current_device() == 0
        assert _relmax(xg.grad, xr.grad) < 1e-5
        assert _relmax(blk.ffn[0].weight.grad, ref.ffn0.weight.grad) < 1e-5


@pytest.mark.parametrize("cin,cout,bias,x_grad", [(256, 256, True, True), (192, 256, True, False), (64, 512, False, True),
                                                  (320, 192, True, True)])
def test_conv1x1_forward_backward(cin, cout, bias, x_grad):
    """training.Conv2d (the models' plain 1x1 convolutions) against torch.autograd in fp64; other configurations stay torch's."""
    torch.manual_seed(cin + cout)
    ref = nn.Conv2d(cin, cout, 1, bias=bias).double()
    conv = T.Conv2d(cin, cout, 1, bias=bias).cuda()
    with torch.no_grad():
        conv.weight.copy_(ref.weight.float())
        if bias:
            conv.bias.copy_(ref.bias.float())
    x = torch.randn(2, cin, 9, 13, dtype=torch.float64)
    gout = torch.randn(2, cout, 9, 13, dtype=torch.float64) * 1e-5
    xr = x.clone().requires_grad_(x_grad)
    ref(xr).backward(gout)
    for rep in range(3):                       # direct call, graph capture, graph replay
        conv.zero_grad()
        xg = x.float().cuda().requires_grad_(x_grad)
        y = conv(xg)
        assert isinstance(y.grad_fn, T._Conv1x1Fn._backward_cls)
        y.backward(gout.float().cuda())
        assert _relmax(y, ref(x)) < 2e-5
        assert _relmax(conv.weight.grad, ref.weight.grad) < 1e-5
        if bias:
            assert _relmax(conv.bias.grad, ref.bias.grad) < 1e-5
        if x_grad:
            assert _relmax(xg.grad, xr.grad) < 1e-5
        else:
            assert xg.grad is None
    for other in (T.Conv2d(32, 32, 3, padding=1, groups=32), T.Conv2d(1, 64, 3, padding=1), T.Conv2d(32, 32, 5, padding=2)):
        other = other.cuda()                                # depthwise / one input channel / 5x5: torch's own path
        xo = torch.randn(1, other.in_channels, 6, 6, device="cuda")
        assert not other._on_engine(xo)
        assert other(xo).shape == (1, other.out_channels, 6, 6)


@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W", [(128, 128, 2, 2, 0, 8, 12),      # hyper / temporal-prior downs
                                                       (256, 128, 3, 2, 1, 10, 14),    # encoder.down, mask_sft.down
                                                       (128, 256, 3, 1, 1, 7, 9),      # the sub-pixel 3x3 of decoder.up
                                                       (32, 64, 3, 2, 1, 9, 11)])      # odd sizes, stride 2
def test_conv_kxk_forward_backward(cin, cout, k, stride, pad, H, W):
    """training.Conv2d's k x k instances (im2col view) against torch.autograd in fp64."""
    torch.manual_seed(cin + cout + k)
    ref = nn.Conv2d(cin, cout, k, stride=stride, padding=pad).double()
    conv = T.Conv2d(cin, cout, k, stride=stride, padding=pad).cuda()
    with torch.no_grad():
        conv.weight.copy_(ref.weight.float())
        conv.bias.copy_(ref.bias.float())
    x = torch.randn(2, cin, H, W, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    yr = ref(xr)
    gout = torch.randn_like(yr) * 1e-4
    yr.backward(gout)
    for rep in range(3):                       # direct call, graph capture, graph replay
        conv.zero_grad()
        xg = x.float().cuda().requires_grad_(True)
        y = conv(xg)
        assert isinstance(y.grad_fn, T._Conv1x1Fn._backward_cls) and y.shape == yr.shape
        y.backward(gout.float().cuda())
        errs = {"out": _relmax(y, yr), "w": _relmax(conv.weight.grad, ref.weight.grad),
                "b": _relmax(conv.bias.grad, ref.bias.grad), "x": _relmax(xg.grad, xr.grad)}
        assert errs["out"] < 2e-5 and max(errs["w"], errs["b"], errs["x"]) < 1e-5, errs
