"""numpy / scipy restatement of the reference's volume augmentations (augment/unet3d_augment/transforms.py) — TEST
INFRASTRUCTURE (see oracle/__init__).  The reference's transforms ARE numpy + scipy.ndimage calls; the restatement keeps
those calls for the trivial ones and spells out the two scipy routines the device kernels re-implement
(``gaussian_filter`` and ``map_coordinates`` with mode='reflect'), checked against scipy itself in the CPU tests.
Pinned against the reference's own classes by tests/golden/augment_r2.npz (oracle/make_golden_r2.py)."""
from __future__ import annotations

import numpy as np


def reflect_index(i, n):
    """scipy mode='reflect' (half-sample symmetric): (d c b a | a b c d | d c b a)."""
    j = np.mod(i, 2 * n)
    return np.where(j >= n, 2 * n - 1 - j, j)


def random_flip(m, rs, axis_prob=0.5):
    """transforms.py:39-50."""
    for axis in (0, 1, 2):
        if rs.uniform() > axis_prob:
            m = np.flip(m, axis) if m.ndim == 3 else np.stack([np.flip(c, axis) for c in m], axis=0)
    return m


def random_rotate90(m, rs):
    """transforms.py:68-80."""
    k = rs.randint(0, 4)
    return np.rot90(m, k, (1, 2)) if m.ndim == 3 else np.stack([np.rot90(c, k, (1, 2)) for c in m], axis=0)


def random_contrast(m, rs, alpha=(0.5, 1.5), mean=0.0, execution_probability=0.1):
    """transforms.py:128-135."""
    if rs.uniform() < execution_probability:
        a = rs.uniform(alpha[0], alpha[1])
        return np.clip(mean + a * (m - mean), -1, 1)
    return m


def standardize(m, eps=1e-10, channelwise=False):
    """transforms.py:508-523."""
    if channelwise:
        axes = tuple(range(1, m.ndim))
        mean, std = np.mean(m, axis=axes, keepdims=True), np.std(m, axis=axes, keepdims=True)
    else:
        mean, std = np.mean(m), np.std(m)
    return (m - mean) / np.clip(std, a_min=eps, a_max=None)


def normalize(m, norm01=False, channelwise=False, eps=1e-10):
    """transforms.py:565-605 (no fixed min / max)."""
    if channelwise:
        axes = tuple(range(1, m.ndim))
        lo, hi = np.min(m, axis=axes, keepdims=True), np.max(m, axis=axes, keepdims=True)
    else:
        lo, hi = np.min(m), np.max(m)
    n01 = (m - lo) / (hi - lo + eps)
    return np.clip(n01, 0, 1) if norm01 else np.clip(2 * n01 - 1, -1, 1)


def gaussian_filter(x, sigma):
    """scipy.ndimage.gaussian_filter(x, sigma, mode='reflect', truncate=4.0), axis by axis."""
    lw = int(4.0 * sigma + 0.5)
    k = np.arange(-lw, lw + 1)
    wgt = np.exp(-0.5 * k * k / (sigma * sigma))
    wgt /= wgt.sum()
    out = np.asarray(x, dtype=np.float64)
    for axis in range(out.ndim):
        n = out.shape[axis]
        idx = np.arange(n)
        acc = np.zeros_like(out)
        for kk, wk in zip(k, wgt):
            acc += wk * np.take(out, reflect_index(idx + kk, n), axis=axis)
        out = acc
    return out


def map_coordinates(a, coords, order):
    """scipy.ndimage.map_coordinates(a, coords, order in (0, 1), mode='reflect') of a 3-D array."""
    d, h, w = a.shape
    z, y, x = coords
    if order == 0:
        iz, iy, ix = (np.floor(c + 0.5).astype(np.int64) for c in (z, y, x))
        return a[reflect_index(iz, d), reflect_index(iy, h), reflect_index(ix, w)]
    z0, y0, x0 = (np.floor(c).astype(np.int64) for c in (z, y, x))
    tz, ty, tx = z - z0, y - y0, x - x0
    out = 0.0
    for oz, wz in ((0, 1 - tz), (1, tz)):
        for oy, wy in ((0, 1 - ty), (1, ty)):
            for ox, wx in ((0, 1 - tx), (1, tx)):
                out = out + wz * wy * wx * a[reflect_index(z0 + oz, d), reflect_index(y0 + oy, h),
                                             reflect_index(x0 + ox, w)].astype(np.float64)
    return out.astype(a.dtype)


def elastic_deformation(m, rs, spline_order, alpha=2000, sigma=50, execution_probability=0.1, apply_3d=True):
    """transforms.py:158-191."""
    if rs.uniform() < execution_probability:
        shape = m.shape if m.ndim == 3 else m[0].shape
        dz = gaussian_filter(rs.randn(*shape), sigma) * alpha if apply_3d else np.zeros(shape)
        dy, dx = [gaussian_filter(rs.randn(*shape), sigma) * alpha for _ in range(2)]
        zz, yy, xx = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), np.arange(shape[2]), indexing="ij")
        idx = zz + dz, yy + dy, xx + dx
        if m.ndim == 3:
            return map_coordinates(m, idx, spline_order)
        return np.stack([map_coordinates(c, idx, spline_order) for c in m], axis=0)
    return m
