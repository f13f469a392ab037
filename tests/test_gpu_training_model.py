"""Training step of the reference's OWN `performance` model with the engine's blocks swapped in.

oracle/_ref holds the unmodified reference modules.  The same class is built twice -- stock, and inside
`training.reference_patched` (DepthConvBlock, AdaptiveQuant from the engine; likelihood through `training.adopt`) --,
given the same parameters, the same inputs and the same generator state for the noise quantiser, and stepped through the
trainer's loss (trainer_seg_video_model.py:904-934: bpp_y + bpp_z + lambda * mse) in train mode, fp32 with TF32 off.
Compared: loss terms and the gradient of EVERY parameter.  The STE rounding makes the forward discontinuous: an
activation that lands on the other side of .5 in one of the two runs moves the gradients by a finite amount, so the
gradient gate is looser than the block-level one (tests/test_gpu_training.py) and the measured values are printed.

Two configurations.  (a) blocks + quantisers from the engine, likelihood left to the reference's torch code: every
difference comes from the engine's blocks -- this one carries the gradient gate.  (b) likelihood through the engine too
(`adopt`): its forward reproduces the CPU reference's erf (MKL saturation point), the stock model here
runs torch's CUDA erf, and for tail symbols (p ~ 1e-8, 0.1 % of the elements with random-init weights) the fp32
autograd gradient -1 / (p ln 2) * dp is dominated by the cancellation noise of p itself, while the engine evaluates the
derivative in fp64.  Those elements carry large gradients, so (b) is held to the bpp gate on the loss terms and its
gradient deviation is only reported; the likelihood gradient itself is pinned element-wise, away from the
cancellation band, in tests/test_gpu_training.py::test_gaussian_bits_backward.
"""
import sys

import pytest
import torch
import torch.nn.functional as F

from helpers import D

pytestmark = pytest.mark.gpu
T = D.training


def _reference():
    from oracle import make_ref
    if not make_ref.available():
        pytest.skip("oracle/_ref not present")
    return make_ref.load_reference()


def _step(model, x, qp, dpb, after_i, target, seed):
    model.zero_grad(set_to_none=True)
    torch.manual_seed(seed)
    r = model(x, qp, dpb, after_i=after_i)
    bpp_y, bpp_z = r["bpp_y"].mean(), r["bpp_z"].mean()
    mse = F.mse_loss(r["dpb"]["frame"], target)
    loss = bpp_y + bpp_z + 256.0 * mse
    loss.backward()
    return r, {"loss": loss.item(), "bpp_y": bpp_y.item(), "bpp_z": bpp_z.item(), "mse": mse.item()}


@pytest.mark.parametrize("engine_likelihood", [False, True], ids=["blocks+quant", "blocks+quant+likelihood"])
def test_reference_performance_model_trains_on_engine_blocks(engine_likelihood):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    R = _reference()
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    stock = R["performance"]().to(dev).train()
    mods = [sys.modules[n] for n in ("src.layers.layers", "src.refactor.common_model", "src.refactor.seg_video_model")]
    with T.reference_patched(*mods):
        ours = R["performance"]().to(dev).train()
    if engine_likelihood:
        T.adopt(ours, formula=1)
    assert any(isinstance(m, T.DepthConvBlock) for m in ours.modules())
    assert not any(isinstance(m, T.DepthConvBlock) for m in stock.modules())
    ours.load_state_dict(stock.state_dict())            # same names, same shapes

    H, W = 128, 192
    frames, masks = D.clips.synthetic_clip(3, 1, 3, H, W)
    x = torch.cat([frames, masks], 2).to(dev)
    qp = 32
    stats = {}
    dpb_s = dpb_o = {"frame": x[:, 0, :3].contiguous(), "feature": None}
    for t, after_i in ((1, True), (2, False)):
        rs, ls = _step(stock, x[:, t], qp, dpb_s, after_i, x[:, t, :3], seed=100 + t)
        ro, lo = _step(ours, x[:, t], qp, dpb_o, after_i, x[:, t, :3], seed=100 + t)
        for k in ls:
            rel = abs(lo[k] - ls[k]) / max(abs(ls[k]), 1e-12)
            stats[f"frame{t}.{k}"] = rel
            assert rel < 1e-3, (t, k, lo[k], ls[k])           # the bpp gate of the inference path
        worst, worst_name, n = 0.0, "", 0
        num = den_all = 0.0
        gs = dict(stock.named_parameters())
        for name, p in ours.named_parameters():
            g_ref = gs[name].grad
            if g_ref is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
                continue
            assert p.grad is not None, f"{name}: no gradient through the engine path"
            den = float(g_ref.abs().max())
            if den == 0.0:
                continue
            e = float((p.grad - g_ref).abs().max()) / den
            num += float((p.grad.double() - g_ref.double()).pow(2).sum())
            den_all += float(g_ref.double().pow(2).sum())
            n += 1
            if e > worst:
                worst, worst_name = e, name
        stats[f"frame{t}.worst_grad_rel"] = worst
        l2 = (num / max(den_all, 1e-300)) ** 0.5
        print(f"\nframe {t}: loss terms rel {[f'{stats[f'frame{t}.{k}']:.1e}' for k in ls]}, {n} parameter gradients, "
              f"all gradients as one vector: relative L2 error {l2:.2e}; worst single tensor {worst:.2e} of its max "
              f"({worst_name})")
        if not engine_likelihood:
            # measured on a B200: 3e-4 / 4e-3 (L2), 2.6e-3 / 6e-2 (worst tensor: a bias whose gradient is a small
            # difference of large sums -- the stock fp32 run carries the same kind of error against fp64)
            assert l2 < 2e-2, l2
            assert worst < 0.2, (worst, worst_name)
        dpb_s = {k: v.detach() for k, v in rs["dpb"].items()}
        dpb_o = {k: v.detach() for k, v in ro["dpb"].items()}
    T.release_handles()
