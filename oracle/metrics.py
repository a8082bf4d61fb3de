"""numpy restatement of trainer/metrcis.py (TEST INFRASTRUCTURE, see oracle/__init__)."""
from __future__ import annotations

import numpy as np


def compute_iou(preds: np.ndarray, labels: np.ndarray, threshold: float = 0.5) -> float:
    """trainer/metrcis.py:61-81 — per-sample |P&L| / max(|P|+|L|-|P&L|, 1e-6), batch mean."""
    p = (preds > threshold).astype(np.float32)
    l = (labels > threshold).astype(np.float32)
    axes = tuple(range(1, p.ndim))
    inter = (p * l).sum(axis=axes)
    union = np.maximum(p.sum(axis=axes) + l.sum(axis=axes) - inter, 1e-6)
    return float(np.mean(inter / union))


def compute_dice(preds: np.ndarray, labels: np.ndarray, threshold: float = 0.5) -> float:
    """trainer/metrcis.py:84-109 — (2|P&L| + e) / ((|P| + e) + (|L| + e)), e = 1e-6, batch mean."""
    p = (preds > threshold).astype(np.float32)
    l = (labels > threshold).astype(np.float32)
    axes = tuple(range(1, p.ndim))
    inter = (p * l).sum(axis=axes)
    eps = 1e-6
    return float(np.mean((2.0 * inter + eps) / ((p.sum(axis=axes) + eps) + (l.sum(axis=axes) + eps))))


def compute_metrics(logits: np.ndarray, labels: np.ndarray) -> dict:
    """trainer/metrcis.py:153-168 without the matplotlib side effect: squeeze C, p = 1/(1+e^-x+1e-6),
    threshold = global mean of p (applied to predictions AND labels)."""
    preds = np.squeeze(logits, axis=1).astype(np.float32)
    labels = np.squeeze(labels, axis=1).astype(np.float32)
    preds = 1 / (1 + np.exp(-preds) + 1e-6)
    thr = np.mean(preds)
    return {"iou": compute_iou(preds, labels, thr), "dice": compute_dice(preds, labels, thr)}


def iou_score(logits: np.ndarray, target: np.ndarray):
    """trainer/metrcis.py:127-140 (numpy in, sigmoid applied here as the reference does for tensors)."""
    out = 1.0 / (1.0 + np.exp(-logits.astype(np.float64)))
    o, t = out > 0.5, target > 0.5
    smooth = 1e-5
    iou = ((o & t).sum() + smooth) / ((o | t).sum() + smooth)
    return iou, (2 * iou) / (iou + 1)


def dice_coef(logits: np.ndarray, target: np.ndarray) -> float:
    """trainer/metrcis.py:143-150."""
    out = (1.0 / (1.0 + np.exp(-logits.astype(np.float32)))).reshape(-1)
    t = target.reshape(-1)
    smooth = 1e-5
    return float((2.0 * (out * t).sum() + smooth) / (out.sum() + t.sum() + smooth))
