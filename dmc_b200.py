"""Import alias: `import dmc_b200` == the package in
semantic-segmentation-guided-neural-video-compression_b200/ (whose name is not a Python identifier)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("semantic-segmentation-guided-neural-video-compression_b200")
sys.modules[__name__] = _pkg
