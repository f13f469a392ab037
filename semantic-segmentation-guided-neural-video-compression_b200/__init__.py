"""B200-native (sm_100a) implementation of the DMC P-frame / DMCI forward pass.

Drop-in for the codec modules of Hrshed/Semantic-Segmentation-Guided-Neural-Video-Compression:
the `nn.Module`s in `modules` keep the reference's constructors, state_dict layout and
`forward(x, qp, dpb, after_i)` signature, and run on hand-written CUDA kernels behind the
C ABI in include/dmc_b200.h.  The directory name is not an identifier; import it with
`importlib.import_module(...)` or through the `dmc_b200` alias module at the repo root.
"""
from .modules import (DMCConfig, DMCI, DMC_fast, DMC_mask_prop, DMC_old, DMC_performance, NonFiniteError,
                      P_MODELS, build_p_model)
from . import _capi, bitstream, build, clips, data, entropy, modules, training  # noqa: F401

__all__ = ["DMCConfig", "DMCI", "DMC_old", "DMC_performance", "DMC_fast", "DMC_mask_prop", "P_MODELS",
           "build_p_model", "NonFiniteError"]
