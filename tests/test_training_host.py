"""Host side of the training-mode blocks (no GPU): the modules keep the reference's parameter layout and refuse to
run without the CUDA path."""
import pytest
import torch

from helpers import D

T = D.training


def _ref_layers():
    """The reference's own layer classes, when its tree is available (not on the GPU box)."""
    import importlib.util
    import os
    path = "/root/reference/src/layers/layers.py"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("ref_layers", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("cin,cout,force", [(48, 64, False), (64, 64, False), (64, 64, True)])
def test_depth_conv_block_state_dict_matches_reference(cin, cout, force):
    ref = _ref_layers().DepthConvBlock(cin, cout, force_adaptor=force)
    blk = T.DepthConvBlock(cin, cout, force_adaptor=force)
    a, b = ref.state_dict(), blk.state_dict()
    assert list(a) == list(b)
    assert all(a[k].shape == b[k].shape for k in a)
    blk.load_state_dict(a)          # a reference checkpoint loads unchanged


def test_training_blocks_refuse_cpu_tensors():
    blk = T.DepthConvBlock(32, 32)
    with pytest.raises(RuntimeError, match="no CPU path"):
        blk(torch.randn(1, 32, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.AdaptiveQuant("noise").train()(torch.randn(4))
    with pytest.raises(RuntimeError, match="no CPU path"):
        T.gaussian_bits(torch.randn(4), torch.rand(4) + 0.1, 1)


def test_adaptive_quant_constructor():
    q = T.AdaptiveQuant(mode="noise", half_bin=0.25)
    assert q.mode == "noise" and q.half_bin == 0.25
    with pytest.raises(AssertionError):
        T.AdaptiveQuant(mode="other")
