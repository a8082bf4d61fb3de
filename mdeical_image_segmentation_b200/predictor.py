"""Sliding-window prediction with everything between the raw volume and the final map kept on the device —
SURVEY.md §8(f) rank 3, the caller right after the network at inference.

Mirrors ``StandardPredictor`` (model/unet3d/predictor.py:65-165) and the patch logic of
dataset/unet3d_dataset/utils.py (``SliceBuilder._build_slices`` / ``_gen_indices`` :85-125, ``mirror_pad``
:314-342, ``remove_padding`` :345-361).  The reference pads the volume on the host, feeds patches through a
``DataLoader``, copies every predicted patch back to the host and accumulates in numpy; here the window
gather (with the mirror padding folded into its index map), the accumulation without halo, the visit count,
the normalisation and the optional arg-max are ``b200_patch_gather / b200_patch_accumulate /
b200_prediction_finalize``; only the final map crosses PCIe.  There is no CPU fallback.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import ops
from .ops import _p, _stream, call


def gen_indices(i: int, k: int, s: int) -> List[int]:
    """``SliceBuilder._gen_indices`` (utils.py:119-125): patch starts 0, s, 2s, ... plus a last patch flush with
    the end of the axis when the stride does not land there."""
    assert i >= k, "Sample size has to be bigger than the patch size"
    out = list(range(0, i - k + 1, s))
    if out[-1] + k < i:
        out.append(i - k)
    return out


def build_slices(volume_shape: Sequence[int], patch_shape: Sequence[int], stride_shape: Sequence[int]):
    """``SliceBuilder._build_slices`` (utils.py:85-117) for a (D, H, W) volume: z-major list of slice triples."""
    (i_z, i_y, i_x), (k_z, k_y, k_x), (s_z, s_y, s_x) = volume_shape, patch_shape, stride_shape
    return [(slice(z, z + k_z), slice(y, y + k_y), slice(x, x + k_x))
            for z in gen_indices(i_z, k_z, s_z) for y in gen_indices(i_y, k_y, s_y)
            for x in gen_indices(i_x, k_x, s_x)]


def _is_2d_model(model: nn.Module) -> bool:
    from .unet3d import UNet2D
    if isinstance(model, nn.DataParallel):
        model = model.module
    return isinstance(model, UNet2D)


class PredictionAccumulator:
    """``prediction_map`` / ``normalization_mask`` of predictor.py:101-104, 147-150 as device buffers."""

    def __init__(self, channels: int, volume_shape: Sequence[int], device):
        self.c, self.shape = channels, tuple(int(v) for v in volume_shape)
        self.map = torch.zeros((channels,) + self.shape, dtype=torch.float32, device=device)
        self.norm = torch.zeros((channels,) + self.shape, dtype=torch.uint8, device=device)

    def add(self, pred: torch.Tensor, origin: Sequence[int], halo: Sequence[int] = (0, 0, 0),
            channel: Optional[int] = None) -> None:
        """pred: one predicted patch (C, pd, ph, pw) INCLUDING its halo; origin: start of the inner patch."""
        ops.require_cuda(pred)
        pred = pred.contiguous().float()
        d, h, w = self.shape
        call("b200_patch_accumulate", self.c, d, h, w, pred.shape[1], pred.shape[2], pred.shape[3], halo[0], halo[1],
             halo[2], _p(pred), pred.shape[0], -1 if channel is None else channel, origin[0], origin[1], origin[2],
             _p(self.map), _p(self.norm), _stream())

    def result(self, save_segmentation: bool = False) -> torch.Tensor:
        """prediction_map / normalization_mask (in place), or its arg-max over channels as int32 (the reference
        stores uint16, predictor.py:163-164)."""
        d, h, w = self.shape
        seg = torch.empty(self.shape, dtype=torch.int32, device=self.map.device) if save_segmentation else None
        call("b200_prediction_finalize", self.c, d * h * w, _p(self.map), _p(self.norm), _p(seg), _stream())
        return seg if save_segmentation else self.map


def gather_patch(volume: torch.Tensor, origin: Sequence[int], patch_shape: Sequence[int],
                 halo: Sequence[int], out: torch.Tensor) -> torch.Tensor:
    """out (C, p + 2 halo) = mirror-padded window of ``volume`` (C, D, H, W) around the patch at ``origin``."""
    c, d, h, w = volume.shape
    call("b200_patch_gather", c, d, h, w, _p(volume), origin[0], origin[1], origin[2], halo[0], halo[1], halo[2],
         patch_shape[0] + 2 * halo[0], patch_shape[1] + 2 * halo[1], patch_shape[2] + 2 * halo[2], _p(out), _stream())
    return out


class StandardPredictor:
    """predictor.py:65-165.  Two entry points:

    * ``predict_volume(raw, patch_shape, stride_shape, halo_shape, batch_size)`` — the whole loop on the device:
      raw (D, H, W) or (C, D, H, W) array / tensor in, probability maps (C, D, H, W) float32 (or the uint16
      segmentation) out as a numpy array;
    * ``__call__(test_loader)`` — the reference's call: any loader that yields ``(input, indices)`` batches over a
      dataset with ``volume_shape()`` and ``halo_shape`` (``AbstractHDF5Dataset``'s protocol); patches are
      accumulated on the device instead of through ``.cpu().numpy()``.  Writes ``<output_dir>/<name>_predictions.h5``
      when ``h5py`` is importable and the dataset has a ``file_path``; always returns the result array.
    """

    def __init__(self, model: nn.Module, output_dir: Optional[str] = None, out_channels: int = 1,
                 output_dataset: str = "predictions", save_segmentation: bool = False,
                 prediction_channel: Optional[int] = None, **kwargs):
        self.model, self.output_dir, self.out_channels = model, output_dir, out_channels
        self.output_dataset, self.save_segmentation = output_dataset, save_segmentation
        self.prediction_channel = prediction_channel

    # ------------------------------------------------------------------ helpers
    def _channels(self) -> int:
        return 1 if self.prediction_channel is not None else self.out_channels

    def _forward(self, batch: torch.Tensor) -> torch.Tensor:
        if _is_2d_model(self.model):                     # predictor.py:123-129
            return torch.unsqueeze(self.model(torch.squeeze(batch, dim=-3)), dim=-3)
        return self.model(batch)

    def _finish(self, acc: PredictionAccumulator) -> np.ndarray:
        res = acc.result(self.save_segmentation)
        out = res.cpu().numpy()                          # the only device -> host copy of the loop
        return out.astype("uint16") if self.save_segmentation else out

    # ------------------------------------------------------------------ whole loop on the device
    def predict_volume(self, raw, patch_shape: Sequence[int], stride_shape: Sequence[int],
                       halo_shape: Sequence[int] = (0, 0, 0), batch_size: int = 1) -> np.ndarray:
        device = next(self.model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("StandardPredictor runs on a CUDA model only (no CPU fallback)")
        vol = torch.as_tensor(np.asarray(raw) if not torch.is_tensor(raw) else raw).float()
        if vol.dim() == 3:
            vol = vol.unsqueeze(0)
        vol = vol.to(device, non_blocking=True).contiguous()
        halo = tuple(int(v) for v in halo_shape)
        slices = build_slices(vol.shape[1:], patch_shape, stride_shape)
        acc = PredictionAccumulator(self._channels(), vol.shape[1:], device)
        win = tuple(p + 2 * hh for p, hh in zip(patch_shape, halo))
        self.model.eval()
        with torch.no_grad():
            for start in range(0, len(slices), batch_size):
                chunk = slices[start:start + batch_size]
                batch = torch.empty((len(chunk), vol.shape[0]) + win, dtype=torch.float32, device=device)
                for b, idx in enumerate(chunk):
                    gather_patch(vol, [s.start for s in idx], patch_shape, halo, batch[b])
                pred = self._forward(batch)
                for b, idx in enumerate(chunk):
                    acc.add(pred[b], [s.start for s in idx], halo, self.prediction_channel)
        return self._finish(acc)

    # ------------------------------------------------------------------ the reference's entry point
    def __call__(self, test_loader: Iterable) -> np.ndarray:
        dataset = test_loader.dataset
        device = next(self.model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("StandardPredictor runs on a CUDA model only (no CPU fallback)")
        volume_shape = tuple(dataset.volume_shape())
        halo = tuple(int(v) for v in (getattr(dataset, "halo_shape", None) or (0, 0, 0)))
        acc = PredictionAccumulator(self._channels(), volume_shape, device)
        self.model.eval()
        with torch.no_grad():
            for inp, indices in test_loader:
                inp = inp.to(device, non_blocking=True)
                pred = self._forward(inp)
                for b, index in enumerate(indices):
                    index = tuple(index)[-3:]             # a leading channel slice is ignored, as predictor.py:146
                    acc.add(pred[b], [s.start for s in index], halo, self.prediction_channel)
        result = self._finish(acc)
        path = getattr(dataset, "file_path", None)
        if self.output_dir is not None and path is not None:
            try:
                import h5py
            except ImportError as e:                     # pragma: no cover
                raise RuntimeError("h5py is needed to write the prediction file; call predict_volume() or pass "
                                   "output_dir=None to get the array instead") from e
            name = os.path.splitext(os.path.basename(str(path)))[0] + "_predictions.h5"
            with h5py.File(os.path.join(self.output_dir, name), "w") as f:
                f.create_dataset(self.output_dataset, data=result, compression="gzip")
        return result


class LazyPredictor(StandardPredictor):
    """predictor.py:171-223 — for volumes whose prediction maps must not be held in (device) memory: the
    network still runs on the device, but ``prediction_map`` / ``normalization_mask`` live in the output store
    (HDF5 datasets when ``h5py`` is importable and ``output_dir`` is given, otherwise ``.npy`` memory maps under
    ``output_dir``, otherwise host arrays) and every predicted patch — its halo already removed on the device by
    ``b200_patch_accumulate`` — is added to the store as soon as it is computed.  ``_save_results`` is the
    reference's: block-wise normalisation over slices of 1/27 of the volume with the visited block's mask reset
    to 1, arg-max segmentation block by block.  Slower than ``StandardPredictor`` by construction (one D2H copy
    per patch), exactly as the reference warns."""

    # ------------------------------------------------------------------ store
    def _allocate_prediction_maps(self, output_shape, name: str = "volume"):
        self._h5 = None
        if self.output_dir is not None:
            os.makedirs(self.output_dir, exist_ok=True)
            try:
                import h5py
                self._h5 = h5py.File(os.path.join(self.output_dir, name + "_predictions.h5"), "w")
                pm = self._h5.create_dataset(self.output_dataset, shape=output_shape, dtype="float32", chunks=True,
                                             compression="gzip")
                nm = self._h5.create_dataset("normalization", shape=output_shape, dtype="uint8", chunks=True,
                                             compression="gzip")
                return pm, nm
            except ImportError:
                base = os.path.join(self.output_dir, name)
                pm = np.lib.format.open_memmap(base + "_predictions.npy", mode="w+", dtype="float32",
                                               shape=tuple(output_shape))
                nm = np.lib.format.open_memmap(base + "_normalization.npy", mode="w+", dtype="uint8",
                                               shape=tuple(output_shape))
                pm[...] = 0
                nm[...] = 0
                return pm, nm
        return np.zeros(output_shape, dtype="float32"), np.zeros(output_shape, dtype="uint8")

    def _save_results(self, prediction_map, normalization_mask):
        z, y, x = prediction_map.shape[1:]
        block = (z // 3, y // 3, x // 3)
        channels = prediction_map.shape[0]
        seg = None
        if self.save_segmentation:
            seg = (self._h5.create_dataset("segmentation", shape=(z, y, x), dtype="uint16", chunks=True,
                                           compression="gzip") if self._h5 is not None
                   else np.zeros((z, y, x), dtype="uint16"))
        for idx in build_slices((z, y, x), block, block):
            index = (slice(0, channels),) + tuple(idx)
            prediction_map[index] /= normalization_mask[index]
            normalization_mask[index] = 1          # never normalise an overlapping block twice
            if seg is not None:
                seg[tuple(idx)] = np.argmax(prediction_map[index], axis=0).astype("uint16")
        result = np.asarray(seg[...] if seg is not None else prediction_map[...])
        if self._h5 is not None:
            del self._h5["normalization"]
            if self.save_segmentation:
                del self._h5[self.output_dataset]
            self._h5.close()
            self._h5 = None
        return result

    def _add_patch(self, pred: torch.Tensor, origin, halo, prediction_map, normalization_mask) -> None:
        """One predicted patch (C, pd, ph, pw) including its halo -> the store (host side accumulate)."""
        inner = tuple(int(s) - 2 * int(hh) for s, hh in zip(pred.shape[1:], halo))
        acc = PredictionAccumulator(self._channels(), inner, pred.device)     # halo removal + channel pick
        acc.add(pred, (0, 0, 0), halo, self.prediction_channel)
        patch = acc.map.cpu().numpy()
        index = (slice(0, patch.shape[0]),) + tuple(slice(int(o), int(o) + n) for o, n in zip(origin, inner))
        prediction_map[index] += patch
        normalization_mask[index] += 1

    # ------------------------------------------------------------------ entry points
    def predict_volume(self, raw, patch_shape: Sequence[int], stride_shape: Sequence[int],
                       halo_shape: Sequence[int] = (0, 0, 0), batch_size: int = 1) -> np.ndarray:
        device = next(self.model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("LazyPredictor runs on a CUDA model only (no CPU fallback)")
        vol = torch.as_tensor(np.asarray(raw) if not torch.is_tensor(raw) else raw).float()
        if vol.dim() == 3:
            vol = vol.unsqueeze(0)
        vol = vol.to(device, non_blocking=True).contiguous()
        halo = tuple(int(v) for v in halo_shape)
        slices = build_slices(vol.shape[1:], patch_shape, stride_shape)
        pm, nm = self._allocate_prediction_maps((self._channels(),) + tuple(vol.shape[1:]))
        win = tuple(p + 2 * hh for p, hh in zip(patch_shape, halo))
        self.model.eval()
        with torch.no_grad():
            for start in range(0, len(slices), batch_size):
                chunk = slices[start:start + batch_size]
                batch = torch.empty((len(chunk), vol.shape[0]) + win, dtype=torch.float32, device=device)
                for b, idx in enumerate(chunk):
                    gather_patch(vol, [s.start for s in idx], patch_shape, halo, batch[b])
                pred = self._forward(batch)
                for b, idx in enumerate(chunk):
                    self._add_patch(pred[b], [s.start for s in idx], halo, pm, nm)
        return self._save_results(pm, nm)

    def __call__(self, test_loader: Iterable) -> np.ndarray:
        dataset = test_loader.dataset
        device = next(self.model.parameters()).device
        if device.type != "cuda":
            raise RuntimeError("LazyPredictor runs on a CUDA model only (no CPU fallback)")
        volume_shape = tuple(dataset.volume_shape())
        halo = tuple(int(v) for v in (getattr(dataset, "halo_shape", None) or (0, 0, 0)))
        path = getattr(dataset, "file_path", None)
        name = os.path.splitext(os.path.basename(str(path)))[0] if path is not None else "volume"
        pm, nm = self._allocate_prediction_maps((self._channels(),) + volume_shape, name)
        self.model.eval()
        with torch.no_grad():
            for inp, indices in test_loader:
                pred = self._forward(inp.to(device, non_blocking=True))
                for b, index in enumerate(indices):
                    index = tuple(index)[-3:]
                    self._add_patch(pred[b], [s.start for s in index], halo, pm, nm)
        return self._save_results(pm, nm)
