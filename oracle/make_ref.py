"""TEST / BASELINE INFRASTRUCTURE ONLY.  Copies the reference's codec modules, UNMODIFIED, from the read-only
checkout into oracle/_ref/ so that they travel to the GPU box (oracle/_ref/ is git-ignored, not gpurun-ignored):

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

What is copied: src/models, src/refactor, src/layers (the files SURVEY.md section 8(a) cites), src/utils/stream_helper.py
(section 8(f) rank 1) plus empty package markers.  Nothing is edited; a manifest with the sha256 of every file is written next to them, and
`load_reference()` verifies it before importing.  Users: bench.py (`--impl reference` CPU arm and the
`gpu_eager_baseline` block: the reference's own nn.Modules timed on the host cores / eagerly on the same B200) and
tests.  The product never imports it.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DMC_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "oracle", "_ref")
PACKAGES = ("src/models", "src/refactor", "src/layers")
FILES = ("src/utils/stream_helper.py",)          # the bit-stream container helpers (tests of bitstream.py)


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def make(verbose=True) -> bool:
    """Returns True when oracle/_ref is in place (copied now or already complete)."""
    if not os.path.isdir(os.path.join(REF, "src")):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {}
    for pkg in PACKAGES:
        out = os.path.join(DST, pkg)
        os.makedirs(out, exist_ok=True)
        for name in sorted(os.listdir(os.path.join(REF, pkg))):
            if not name.endswith(".py"):
                continue
            shutil.copyfile(os.path.join(REF, pkg, name), os.path.join(out, name))
            manifest[f"{pkg}/{name}"] = _sha(os.path.join(out, name))
    for rel in FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), os.path.join(DST, rel))
        manifest[rel] = _sha(os.path.join(DST, rel))
    for d in ("src", "src/utils") + PACKAGES:           # the reference relies on namespace packages; markers keep imports local
        marker = os.path.join(DST, d, "__init__.py")
        if not os.path.exists(marker):
            open(marker, "w").close()
    json.dump({"source": REF, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if verbose:
        print(f"oracle/_ref: {len(manifest)} reference files copied unmodified from {REF}")
    return True


def available() -> bool:
    return os.path.exists(os.path.join(DST, "MANIFEST.json"))


def load_reference():
    """Imports the copied reference modules (verifying the manifest) and returns the model classes.
    The copy is imported under its own top-level package name `src`, exactly like the reference's trainer does."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/make_ref.py` where /root/reference exists")
    man = json.load(open(os.path.join(DST, "MANIFEST.json")))
    for rel, digest in man["files"].items():
        if _sha(os.path.join(DST, rel)) != digest:
            raise RuntimeError(f"oracle/_ref/{rel} does not match its manifest (the copy must stay unmodified)")
    sys.dont_write_bytecode = True
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from src.models.image_model import DMCI
    from src.models.video_model import DMC as DMC_old
    from src.refactor.config import DMCConfig
    from src.refactor.mask_prop_seg_video_model import DMC as DMC_mask_prop
    from src.refactor.seg_video_model import DMC as DMC_performance
    from src.refactor.seg_video_model_fast import DMC as DMC_fast
    return {"DMCI": DMCI, "DMCConfig": DMCConfig, "old": lambda: DMC_old(),
            "performance": lambda: DMC_performance(DMCConfig()), "fast": lambda: DMC_fast(DMCConfig()),
            "mask_prop": lambda: DMC_mask_prop(DMCConfig())}


if __name__ == "__main__":
    ok = make()
    sys.exit(0 if ok else 1)
