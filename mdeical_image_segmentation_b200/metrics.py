"""Evaluation metrics of the path, reduced on the device.

Mirrors trainer/metrcis.py (compute_iou :61-81, compute_dice :84-109, iou_score :127-140, dice_coef
:143-150, compute_metrics :153-168) and model/unet3d/metrics.py (DiceCoefficient :15-29, MeanIoU
:32-104).  The reference pulls the full logits to the host and makes ~10 numpy passes; here two
kernels produce (sum of probabilities) and per-sample {|P&L|, |P|, |L|} counts, and only those few
numbers cross PCIe.  numpy inputs (what ``transformers.Trainer`` hands to ``compute_metrics``) are
accepted and staged to the current CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .losses import compute_per_channel_dice


def _dev(a) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        t = a.detach()
        return t.float().cuda() if not t.is_cuda else t.float()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _counts(preds, labels, thr: torch.Tensor, mode: int) -> np.ndarray:
    p, l = _dev(preds), _dev(labels)
    if p.shape != l.shape:
        raise ValueError(f"preds {tuple(p.shape)} and labels {tuple(l.shape)} differ in shape")
    if p.dim() == 1:
        p, l = p[None], l[None]
    return ops.metric_counts(p, l, thr, mode).cpu().numpy()


def _thr_tensor(threshold) -> torch.Tensor:
    if isinstance(threshold, torch.Tensor):
        return threshold.detach().float().reshape(1).cuda()
    return torch.tensor([float(threshold)], dtype=torch.float32, device="cuda")


def _iou_from_counts(c: np.ndarray) -> float:
    inter = c[:, 0].astype(np.float32)
    union = np.maximum(c[:, 1].astype(np.float32) + c[:, 2].astype(np.float32) - inter, 1e-6)
    return float(np.mean(inter / union))


def _dice_from_counts(c: np.ndarray) -> float:
    eps = np.float32(1e-6)
    inter = c[:, 0].astype(np.float32)
    num = 2.0 * inter + eps
    den = (c[:, 1].astype(np.float32) + eps) + (c[:, 2].astype(np.float32) + eps)
    return float(np.mean(num / den))


def compute_iou(preds, labels, threshold=0.5):
    """trainer/metrcis.py:61-81 — preds are probability maps (N, H, W)."""
    return _iou_from_counts(_counts(preds, labels, _thr_tensor(threshold), 2))


def compute_dice(preds, labels, threshold=0.5):
    """trainer/metrcis.py:84-109."""
    return _dice_from_counts(_counts(preds, labels, _thr_tensor(threshold), 2))


def compute_metrics(p):
    """trainer/metrcis.py:153-168 — ``p`` is an EvalPrediction (or any (predictions, label_ids)
    pair) of logits / labels shaped (N, 1, H, W).  The matplotlib side effect (:164) is not part of
    the path and is not reproduced."""
    logits, labels = (p.predictions, p.label_ids) if hasattr(p, "predictions") else p
    x, l = _dev(logits), _dev(labels)
    if x.shape[1] != 1:
        raise ValueError("compute_metrics squeezes axis 1: only single-channel logits are valid")
    x, l = x.squeeze(1).contiguous(), l.squeeze(1).contiguous()
    thr = (ops.metric_prob_sum(x) / x.numel()).float()     # global mean of 1/(1+exp(-x)+1e-6)
    c = ops.metric_counts(x, l, thr, 0).cpu().numpy()
    return {"iou": _iou_from_counts(c), "dice": _dice_from_counts(c)}


def iou_score(output, target):
    """trainer/metrcis.py:127-140 — sigmoid > 0.5 vs target > 0.5, global, smooth 1e-5."""
    x, t = _dev(output).reshape(1, -1), _dev(target).reshape(1, -1)
    c = ops.metric_counts(x, t, _thr_tensor(0.5), 1).cpu().numpy()[0]
    smooth = 1e-5
    inter, union = c[0], c[1] + c[2] - c[0]
    iou = (inter + smooth) / (union + smooth)
    return iou, (2 * iou) / (iou + 1)


def dice_coef(output, target):
    """trainer/metrcis.py:143-150 — soft dice over the flattened tensors."""
    x, t = _dev(output).reshape(1, 1, -1), _dev(target).reshape(1, 1, -1)
    s = ops.seg_sums(x, t, 0).cpu().numpy()[0]
    smooth = 1e-5
    return float((2.0 * s[0] + smooth) / (s[3] + s[4] + smooth))


class DiceCoefficient:
    """model/unet3d/metrics.py:15-29 — mean per-channel dice of probability maps."""

    def __init__(self, epsilon=1e-6, **kwargs):
        self.epsilon = epsilon

    def __call__(self, input, target):
        with torch.no_grad():
            return torch.mean(compute_per_channel_dice(input, target, epsilon=self.epsilon))


class MeanIoU:
    """model/unet3d/metrics.py:32-104 (5-D one-hot targets, ignore_index unsupported)."""

    def __init__(self, skip_channels=(), ignore_index=None, **kwargs):
        if ignore_index is not None:
            raise NotImplementedError("ignore_index is outside the B200 hot path")
        self.skip_channels = tuple(skip_channels)

    def __call__(self, input, target):
        assert input.dim() == 5
        assert input.size() == target.size()
        n, c = input.shape[0], input.shape[1]
        if c == 1:
            cnt = ops.metric_counts(_dev(input).reshape(n, -1), _dev(target).reshape(n, -1),
                                    _thr_tensor(0.5), 2)
            # target.byte() in the reference: compare with > 0.5 is identical for {0,1} masks
            inter, union = cnt[:, 0], cnt[:, 1] + cnt[:, 2] - cnt[:, 0]
            per_sample = (inter / union.clamp(min=1e-8)).float()
            return per_sample.mean().cpu()
        cnt = ops.metric_argmax_counts(_dev(input), _dev(target))
        keep = [k for k in range(c) if k not in self.skip_channels]
        assert keep, "All channels were ignored from the computation"
        iou = (cnt[:, keep, 0] / cnt[:, keep, 1].clamp(min=1e-8)).float()
        return iou.mean(dim=1).mean().cpu()
