"""Recipe for ``oracle/_ref/``: a verbatim copy of the reference files of the U-Net hot path (SURVEY.md
§8a), made from ``/root/reference`` in the build container so that ``bench.py --impl reference`` can run
the reference's OWN modules on the GPU box's host cores (``cpu_baseline.kind == "reference"``).

    python -m oracle.build_ref            # also run by __graft_entry__.build()

``oracle/_ref/`` is build output: listed in ``.gitignore`` (reference sources never enter the history) but
not in ``.gpurunignore`` (it travels with the snapshot like the built ``.so``).  Nothing is edited — the
stale import paths of the reference are satisfied by the shims of ``oracle/ref_loader.py`` at import time.
TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__): the product package never imports it.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("B200_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")

# the path's files only: model/unet2d (networks, layers, losses), model/unet3d (building blocks, networks,
# losses, metrics, SE, HF wrapper, predictor), trainer/metrcis.py (numpy metrics)
FILES = [
    "model/__init__.py",
    "model/unet2d/__init__.py", "model/unet2d/unet.py", "model/unet2d/layers.py", "model/unet2d/loss.py",
    "model/unet2d/init_weights.py",
    "model/unet3d/__init__.py", "model/unet3d/buildingblocks.py", "model/unet3d/model.py",
    "model/unet3d/losses.py", "model/unet3d/metrics.py", "model/unet3d/se.py", "model/unet3d/utils.py",
    "model/unet3d/UNet3D.py", "model/unet3d/predictor.py", "model/unet3d/seg_metrics.py",
    "trainer/__init__.py", "trainer/metrcis.py", "trainer/MYtrainer.py", "trainer/callbacks.py",
]


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref/ is in place (copied now or already there)."""
    if not os.path.isdir(os.path.join(SRC, "model", "unet2d")):
        return os.path.isdir(os.path.join(DST, "model", "unet2d"))
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Verbatim copies of reference files made by oracle/build_ref.py (build output, git-ignored).\n")
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files copied from {SRC}")
    return True


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
