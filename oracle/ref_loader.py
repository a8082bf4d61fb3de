"""Imports the UNMODIFIED reference modules with the import shims its stale import paths need (SURVEY.md
Appendix C) — from /root/reference inside the build container (oracle/make_golden.py, the CPU-side oracle
pin test) or from oracle/_ref, the verbatim copy of the path's files that oracle/build_ref.py makes and that
travels to the GPU box (bench.py's CPU arm only).
TEST INFRASTRUCTURE (see oracle/__init__)."""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

_loaded = None
REF_ROOT = os.environ.get("B200_REFERENCE_ROOT", "/root/reference")
REF_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "unet2d"))


def use_copy() -> bool:
    """Point the loader at oracle/_ref (must be called before load()); False when the copy is absent."""
    global REF_ROOT
    if _loaded is not None:
        return REF_ROOT == REF_COPY
    if not os.path.isdir(os.path.join(REF_COPY, "model", "unet2d")):
        return False
    REF_ROOT = REF_COPY
    return True


def _stub(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference modules: U2 (unet2d.unet), L2 (unet2d.layers),
    LOSS2 (unet2d.loss), M3 (unet3d.model), B3 (unet3d.buildingblocks), L3 (unet3d.losses),
    MET3 (unet3d.metrics), MX (trainer.metrcis)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    class _MS_SSIM(torch.nn.Module):  # only constructed by SegmentationLoss, never called on the path
        def __init__(self, *a, **k):
            super().__init__()

    for name in ("pytorch_msssim", "h5py", "medpy", "medpy.metric", "matplotlib", "matplotlib.pyplot",
                 "skimage", "skimage.measure", "skimage.metrics"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    sys.modules["pytorch_msssim"].__dict__.setdefault("MS_SSIM", _MS_SSIM)
    sys.modules["pytorch_msssim"].__dict__.setdefault("ms_ssim", None)
    sys.modules["matplotlib"].__dict__.setdefault("pyplot", sys.modules["matplotlib.pyplot"])
    sys.modules["medpy"].__dict__.setdefault("metric", sys.modules["medpy.metric"])
    sys.modules["skimage"].__dict__.setdefault("measure", sys.modules["skimage.measure"])
    for n in ("adapted_rand_error", "peak_signal_noise_ratio", "mean_squared_error"):
        sys.modules["skimage.metrics"].__dict__.setdefault(n, None)

    importlib.import_module("model")
    se = importlib.import_module("model.unet3d.se")
    _stub("pytorch3dunet")
    _stub("pytorch3dunet.unet3d")
    sys.modules["pytorch3dunet.unet3d.se"] = se
    ns = types.SimpleNamespace()
    ns.U2 = importlib.import_module("model.unet2d.unet")
    ns.L2 = importlib.import_module("model.unet2d.layers")
    ns.LOSS2 = importlib.import_module("model.unet2d.loss")
    ns.B3 = importlib.import_module("model.unet3d.buildingblocks")
    ns.M3 = importlib.import_module("model.unet3d.model")
    ns.L3 = importlib.import_module("model.unet3d.losses")
    # metrics.py imports pytorch3dunet.unet3d.{losses,seg_metrics,utils}: alias to the local files
    sys.modules["pytorch3dunet.unet3d.losses"] = ns.L3
    _stub("pytorch3dunet.unet3d.seg_metrics", AveragePrecision=None, Accuracy=None)
    u3 = importlib.import_module("model.unet3d.utils")
    _stub("pytorch3dunet.unet3d.utils", get_logger=u3.get_logger,
          expand_as_one_hot=u3.expand_as_one_hot, convert_to_numpy=getattr(u3, "convert_to_numpy", None))
    ns.MET3 = importlib.import_module("model.unet3d.metrics")
    importlib.import_module("trainer")
    ns.MX = importlib.import_module("trainer.metrcis")
    _loaded = ns
    return ns
