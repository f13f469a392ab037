"""Mints tests/golden/data_path.npz from the REFERENCE's own colour conversion (run in the build container, where
/root/reference exists):  python oracle/make_golden_data.py

`src/dataset/seg_waymo_dataset.py` imports cv2 and the Waymo reader at module level (both absent here), so the two
functions are taken from its source with `ast` and executed unmodified; `cv2.imdecode` / `cvtColor` are replaced by
handing `_rgb_from_proto`'s last line an already decoded RGB array (decoding is not part of the path).
"""
import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/dataset/seg_waymo_dataset.py"


def reference_functions():
    src = open(REF).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "_rgb_to_ycbcr_bt709":
            exec(compile(ast.Module([node], []), REF, "exec"), ns)          # noqa: S102
        if isinstance(node, ast.FunctionDef) and node.name == "_rgb_from_proto":
            # its last statement is the uint8 -> [0,1] float conversion; the lines before it decode the JPEG
            last = ast.get_source_segment(src, node.body[-1])
            assert last.strip().startswith("return torch.as_tensor(rgb"), last
            ns["_rgb_to_float_src"] = last.strip()[len("return "):]
    return ns


def main():
    ns = reference_functions()
    g = np.random.default_rng(20261018)
    T, H, W = 2, 40, 52
    img = g.integers(0, 256, size=(T, H, W, 3), dtype=np.uint8)
    img[0, 0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0], [0, 255, 255],
                     [255, 0, 255]]                                        # the corners of the cube: exercises the clamp
    mask = (g.random((T, H, W)) < 0.3).astype(np.uint8)
    out = []
    for t in range(T):
        rgb = img[t]
        chw = eval(ns["_rgb_to_float_src"], {"torch": torch, "rgb": rgb})   # noqa: S307  (the reference's own expression)
        out.append(ns["_rgb_to_ycbcr_bt709"](chw).numpy())
    path = os.path.join(ROOT, "tests", "golden", "data_path.npz")
    np.savez_compressed(path, img=img, mask=mask, ycbcr=np.stack(out), rgb_to_float=ns["_rgb_to_float_src"])
    print("wrote", path, np.stack(out).shape, "expression:", ns["_rgb_to_float_src"])


if __name__ == "__main__":
    sys.dont_write_bytecode = True
    main()
