"""GPU-side volume augmentation — SURVEY.md §8(f) rank 5: the reference's numpy / scipy transforms
(augment/unet3d_augment/transforms.py) with the same class names, constructor arguments and random-decision
order, executed by ``csrc/augment.cu`` on CUDA tensors (3-D ``D x H x W`` or 4-D ``C x D x H x W`` float32).

The random decisions (flip per axis, number of rotations, contrast alpha, execution probabilities) are drawn from the
caller's ``np.random.RandomState`` in the reference's order, so raw and label pipelines seeded alike stay consistent
and a pipeline reproduces the reference's choices for the same seed.  ``ElasticDeformation`` draws its displacement
noise on the device by default (``noise="device"``; three (D, H, W) normal fields per call are what made the
reference's version cost ~1 s per patch on the host); ``noise="host"`` draws them from the ``RandomState`` exactly as
the reference does (used by the parity tests).  Spline orders 0 and 1 are implemented (labels / raw); the label
conversion transforms (boundaries, affinities) stay on the host.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from .ops import _p, _stream, call, require_cuda


def _as4(m: torch.Tensor):
    require_cuda(m)
    assert m.dim() in (3, 4), "Supports only 3D (DxHxW) or 4D (CxDxHxW) images"
    x = m.contiguous().float()
    return (x.unsqueeze(0), True) if x.dim() == 3 else (x, False)


def _back(x: torch.Tensor, squeezed: bool) -> torch.Tensor:
    return x[0] if squeezed else x


def _affine(x: torch.Tensor, a, b, lo=float("-inf"), hi=float("inf")) -> torch.Tensor:
    c = x.shape[0]
    a = torch.as_tensor(a, dtype=torch.float32, device=x.device).reshape(-1).expand(c).contiguous()
    b = torch.as_tensor(b, dtype=torch.float32, device=x.device).reshape(-1).expand(c).contiguous()
    out = torch.empty_like(x)
    call("b200_aug_affine_clip", c, x[0].numel(), _p(x), _p(a), _p(b), float(lo), float(hi), _p(out), _stream())
    return out


def channel_stats(x: torch.Tensor) -> torch.Tensor:
    """(C, 4) fp64: sum, sum of squares, min, max per channel of a (C, D, H, W) tensor."""
    c = x.shape[0]
    stats = torch.tensor([0.0, 0.0, float("inf"), float("-inf")], dtype=torch.float64, device=x.device).repeat(c, 1)
    call("b200_aug_channel_stats", c, x[0].numel(), _p(x), _p(stats), _stream())
    return stats


class Compose:
    """transforms.py:15-22."""

    def __init__(self, transforms):
        self.transforms = transforms

    def __call__(self, m):
        for t in self.transforms:
            m = t(m)
        return m


class RandomFlip:
    """transforms.py:25-50 — every axis is flipped when ``random_state.uniform() > axis_prob``."""

    def __init__(self, random_state, axis_prob=0.5, **kwargs):
        assert random_state is not None, "RandomState cannot be None"
        self.random_state, self.axes, self.axis_prob = random_state, (0, 1, 2), axis_prob

    def __call__(self, m):
        x, sq = _as4(m)
        bits = 0
        for axis in self.axes:
            if self.random_state.uniform() > self.axis_prob:
                bits |= 1 << axis
        if bits == 0:
            return m
        out = torch.empty_like(x)
        c, d, h, w = x.shape
        call("b200_aug_flip_rot90", c, d, h, w, bits, 0, _p(x), _p(out), _stream())
        return _back(out, sq)


class RandomRotate90:
    """transforms.py:53-80 — ``k = random_state.randint(0, 4)`` quarter turns in the (H, W) plane."""

    def __init__(self, random_state, **kwargs):
        self.random_state, self.axis = random_state, (1, 2)

    def __call__(self, m):
        x, sq = _as4(m)
        k = int(self.random_state.randint(0, 4))
        if k == 0:
            return m
        c, d, h, w = x.shape
        out = torch.empty((c, d, w, h) if k & 1 else (c, d, h, w), dtype=torch.float32, device=x.device)
        call("b200_aug_flip_rot90", c, d, h, w, 0, k, _p(x), _p(out), _stream())
        return _back(out, sq)


class RandomContrast:
    """transforms.py:115-135 — ``clip(mean + alpha (v - mean), -1, 1)`` with probability ``execution_probability``."""

    def __init__(self, random_state, alpha=(0.5, 1.5), mean=0.0, execution_probability=0.1, **kwargs):
        assert len(alpha) == 2
        self.random_state, self.alpha, self.mean = random_state, alpha, mean
        self.execution_probability = execution_probability

    def __call__(self, m):
        if self.random_state.uniform() < self.execution_probability:
            alpha = self.random_state.uniform(self.alpha[0], self.alpha[1])
            x, sq = _as4(m)
            return _back(_affine(x, alpha, self.mean - alpha * self.mean, -1.0, 1.0), sq)
        return m


class Standardize:
    """transforms.py:495-523 — z-score with given or computed (global / channel-wise) mean and population std."""

    def __init__(self, eps=1e-10, mean=None, std=None, channelwise=False, **kwargs):
        if mean is not None or std is not None:
            assert mean is not None and std is not None
        self.mean, self.std, self.eps, self.channelwise = mean, std, eps, channelwise

    def __call__(self, m):
        x, sq = _as4(m)
        if self.mean is not None:
            mean = torch.as_tensor(self.mean, dtype=torch.float64, device=x.device).reshape(-1)
            std = torch.as_tensor(self.std, dtype=torch.float64, device=x.device).reshape(-1)
        else:
            st = channel_stats(x)
            n = float(x[0].numel())
            if self.channelwise and not sq:
                mean = st[:, 0] / n
                std = (st[:, 1] / n - mean * mean).clamp_(min=0).sqrt()
            else:
                tot = n * x.shape[0]
                mean = (st[:, 0].sum() / tot).reshape(1)
                std = (st[:, 1].sum() / tot - mean * mean).clamp_(min=0).sqrt()
        inv = 1.0 / std.clamp(min=self.eps)
        return _back(_affine(x, inv.float(), (-mean * inv).float()), sq)


class Normalize:
    """transforms.py:547-605 — min-max scaling to [-1, 1] (or [0, 1] with ``norm01``), global or channel-wise, with
    optional fixed ``min_value`` / ``max_value``."""

    def __init__(self, min_value=None, max_value=None, norm01=False, channelwise=False, eps=1e-10, **kwargs):
        if min_value is not None and max_value is not None and not channelwise:
            assert max_value > min_value
        self.min_value, self.max_value, self.norm01 = min_value, max_value, norm01
        self.channelwise, self.eps = channelwise, eps

    def __call__(self, m):
        x, sq = _as4(m)
        st = channel_stats(x)
        if self.channelwise and not sq:
            lo, hi = st[:, 2].clone(), st[:, 3].clone()
            for src, dst in ((self.min_value, lo), (self.max_value, hi)):
                if src is not None:
                    for i, v in enumerate(src):
                        if v != "None":
                            dst[i] = float(v)
        else:
            lo = st[:, 2].min().reshape(1) if self.min_value is None else torch.tensor([float(self.min_value)],
                                                                                      dtype=torch.float64,
                                                                                      device=x.device)
            hi = st[:, 3].max().reshape(1) if self.max_value is None else torch.tensor([float(self.max_value)],
                                                                                      dtype=torch.float64,
                                                                                      device=x.device)
        inv = 1.0 / (hi - lo + self.eps)
        if self.norm01:
            return _back(_affine(x, inv.float(), (-lo * inv).float(), 0.0, 1.0), sq)
        return _back(_affine(x, (2 * inv).float(), (-2 * lo * inv - 1).float(), -1.0, 1.0), sq)


class AdditiveGaussianNoise:
    """transforms.py:608-619 — ``m + N(0, std)``, ``std ~ U(scale)``; the noise itself is drawn on the device."""

    def __init__(self, random_state, scale=(0.0, 1.0), execution_probability=0.1, **kwargs):
        self.execution_probability, self.random_state, self.scale = execution_probability, random_state, scale

    def __call__(self, m):
        if self.random_state.uniform() < self.execution_probability:
            std = self.random_state.uniform(self.scale[0], self.scale[1])
            require_cuda(m)
            return m + torch.randn(m.shape, device=m.device, dtype=torch.float32) * float(std)
        return m


def gaussian_filter(field: torch.Tensor, sigma: float) -> torch.Tensor:
    """``scipy.ndimage.gaussian_filter(field, sigma, mode='reflect')`` of a (D, H, W) float32 device tensor."""
    require_cuda(field)
    a = field.contiguous().float()
    d, h, w = a.shape
    b = torch.empty_like(a)
    for axis in range(3):
        call("b200_aug_gaussian_axis", d, h, w, axis, float(sigma), _p(a), _p(b), _stream())
        a, b = b, a
    return a


def map_coordinates(x: torch.Tensor, dz: Optional[torch.Tensor], dy: torch.Tensor, dx: torch.Tensor,
                    order: int) -> torch.Tensor:
    """``scipy.ndimage.map_coordinates(x[c], (z + dz, y + dy, x + dx), order, mode='reflect')`` per channel."""
    c, d, h, w = x.shape
    out = torch.empty_like(x)
    call("b200_aug_map_coordinates", c, d, h, w, _p(x), _p(dz), _p(dy), _p(dx), int(order), _p(out), _stream())
    return out


class ElasticDeformation:
    """transforms.py:138-191 — per-voxel displacement fields ``gaussian_filter(randn, sigma) * alpha`` (three with
    ``apply_3d``), sampled with ``map_coordinates(order=spline_order, mode='reflect')``; orders 0 and 1."""

    def __init__(self, random_state, spline_order, alpha=2000, sigma=50, execution_probability=0.1, apply_3d=True,
                 noise="device", **kwargs):
        if spline_order not in (0, 1):
            raise NotImplementedError("the B200 elastic deformation implements spline orders 0 (labels) and 1")
        self.random_state, self.spline_order, self.alpha, self.sigma = random_state, spline_order, alpha, sigma
        self.execution_probability, self.apply_3d, self.noise = execution_probability, apply_3d, noise

    def _field(self, shape, device):
        if self.noise == "host":      # the reference's stream: random_state.randn(*shape) (transforms.py:166, 172)
            n = torch.from_numpy(self.random_state.randn(*shape).astype(np.float32)).to(device)
        else:
            n = torch.randn(shape, device=device, dtype=torch.float32)
        return gaussian_filter(n, self.sigma) * float(self.alpha)

    def __call__(self, m):
        if self.random_state.uniform() < self.execution_probability:
            x, sq = _as4(m)
            shape = tuple(x.shape[1:])
            dz = self._field(shape, x.device) if self.apply_3d else None
            dy = self._field(shape, x.device)
            dx = self._field(shape, x.device)
            return _back(map_coordinates(x, dz, dy, dx, self.spline_order), sq)
        return m


class ToTensor:
    """transforms.py:636-655 on the device: optional channel axis, dtype cast."""

    def __init__(self, expand_dims, dtype=torch.float32, **kwargs):
        self.expand_dims = expand_dims
        self.dtype = dtype if isinstance(dtype, torch.dtype) else torch.float32

    def __call__(self, m):
        assert m.dim() in (3, 4), "Supports only 3D (DxHxW) or 4D (CxDxHxW) images"
        if self.expand_dims and m.dim() == 3:
            m = m.unsqueeze(0)
        return m.to(self.dtype)
