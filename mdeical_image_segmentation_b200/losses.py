"""Loss modules of the path, all built on ONE fused reduction kernel.

``b200_seg_sums`` reads the fp32 logits and the target once and returns, per channel, the six sums
every loss of the reference is a function of (sum p*t, p*p, t*t, p, t, bce).  The loss value is then
a few scalar ops on that (C, 6) tensor — written with ordinary torch arithmetic so the reference's
formulas (clamps, epsilons) are restated literally — and its gradient w.r.t. the sums is turned back
into d(loss)/d(logits) by one element-wise kernel (``b200_seg_loss_bwd``).

Mirrors: model/unet3d/losses.py (compute_per_channel_dice :7-33, DiceLoss :119-129, BCEDiceLoss
:167-178, get_loss_criterion :273-306), model/unet2d/loss.py (IoULoss :32-42, F1Loss :46-57) and the
criteria picked at model/unet2d/unet.py:1184-1188.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import ops

S_PT, S_PP, S_TT, S_P, S_T, S_BCE = range(6)


class _SegSums(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, target: torch.Tensor, normalization: int):
        ops.require_cuda(x)
        x = x.contiguous().float()
        target = target.contiguous().float()
        ctx.save_for_backward(x, target)
        ctx.normalization = normalization
        return ops.seg_sums(x, target, normalization)

    @staticmethod
    def backward(ctx, gsums):
        x, target = ctx.saved_tensors
        g = gsums.float()
        # d/dp of (sum p t, sum p^2, sum p) and d/dx of sum bce = (p - t)
        coef = torch.stack([g[:, S_BCE], g[:, S_PT], 2.0 * g[:, S_PP], g[:, S_P]], dim=1).contiguous()
        return ops.seg_loss_bwd(x, target, coef, ctx.normalization), None, None


def seg_sums(x: torch.Tensor, target: torch.Tensor, normalization: str = "sigmoid") -> torch.Tensor:
    """(C, 6) fp64 sums over batch and space; differentiable w.r.t. ``x``."""
    if x.shape != target.shape:
        raise ValueError("'input' and 'target' must have the same shape")
    if x.dim() < 3:
        x = x.reshape(1, 1, -1)
        target = target.reshape(1, 1, -1)
    return _SegSums.apply(x, target, {"sigmoid": 0, "none": 1}[normalization])


def compute_per_channel_dice(input, target, epsilon=1e-6, weight=None):
    """losses.py:7-33; ``input`` is a probability map (N, C, ...)."""
    s = seg_sums(input, target, "none")
    intersect = s[:, S_PT]
    if weight is not None:
        intersect = weight.to(intersect) * intersect
    denominator = s[:, S_PP] + s[:, S_TT]
    return (2 * (intersect / denominator.clamp(min=epsilon))).float()


def _dice_from_sums(s, weight, epsilon=1e-6):
    intersect = s[:, S_PT]
    if weight is not None:
        intersect = weight.to(intersect) * intersect
    return 2 * (intersect / (s[:, S_PP] + s[:, S_TT]).clamp(min=epsilon))


class _AbstractDiceLoss(nn.Module):
    def __init__(self, weight=None, normalization="sigmoid"):
        super().__init__()
        self.register_buffer("weight", weight)
        assert normalization in ["sigmoid", "softmax", "none"]
        self.normalization = normalization

    def dice(self, sums, weight):
        raise NotImplementedError

    def forward(self, input, target):
        if self.normalization == "softmax":
            sums = seg_sums(torch.softmax(input, dim=1), target, "none")
        else:
            sums = seg_sums(input, target, self.normalization)
        return (1.0 - torch.mean(self.dice(sums, self.weight))).float()


class DiceLoss(_AbstractDiceLoss):
    """losses.py:119-129 — 1 - mean_c 2*sum(p t) / clamp(sum p^2 + sum t^2, 1e-6)."""

    def dice(self, sums, weight):
        return _dice_from_sums(sums, weight)


class BCEDiceLoss(nn.Module):
    """losses.py:167-178 — alpha * BCEWithLogits + beta * Dice, one pass over logits and target."""

    def __init__(self, alpha, beta):
        super().__init__()
        self.alpha, self.beta = alpha, beta

    def forward(self, input, target):
        s = seg_sums(input, target, "sigmoid")
        bce = s[:, S_BCE].sum() / input.numel()
        dice = 1.0 - torch.mean(_dice_from_sums(s, None))
        return (self.alpha * bce + self.beta * dice).float()


class BCEWithLogitsLoss(nn.Module):
    """nn.BCEWithLogitsLoss() (mean) as chosen at model/unet2d/unet.py:1187."""

    def forward(self, input, target):
        s = seg_sums(input, target.to(input.dtype).reshape(input.shape), "sigmoid")
        return (s[:, S_BCE].sum() / input.numel()).float()


class _CrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        ops.require_cuda(logits)
        if tuple(labels.shape) != (logits.shape[0],) + tuple(logits.shape[2:]):
            raise ValueError(f"CrossEntropyLoss: class-index target of shape {tuple(labels.shape)} does not match "
                             f"logits {tuple(logits.shape)} (expected (N, *spatial), as nn.CrossEntropyLoss)")
        logits = logits.contiguous().float()
        labels = labels.contiguous().long()
        s = ops.ce_fwd(logits, labels)
        # mean over the positions whose label is not ignore_index (-100); a label outside [0, C) poisons the
        # loss with NaN (the stock op raises a device-side assert; an out-of-bounds read must never happen)
        valid = torch.where(s[2] > 0, torch.full_like(s[1], float("nan")), s[1])
        ctx.save_for_backward(logits, labels, valid)
        return (s[0] / valid).float()

    @staticmethod
    def backward(ctx, g):
        logits, labels, valid = ctx.saved_tensors
        scale = (g.double() / valid).float().reshape(1)
        return ops.ce_bwd(logits, labels, scale), None


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() (mean, class-index targets, ignore_index = -100) as chosen at
    model/unet2d/unet.py:1185."""

    def forward(self, input, target):
        return _CrossEntropy.apply(input, target)


class IoULoss(nn.Module):
    """model/unet2d/loss.py:32-42 — global soft IoU over the whole batch tensor."""

    def __init__(self, epsilon=1e-7):
        super().__init__()
        self.epsilon = epsilon

    def forward(self, inputs, targets):
        s = seg_sums(inputs, targets, "sigmoid").sum(dim=0)
        inter = s[S_PT]
        union = s[S_P] + s[S_T] - inter
        return (1 - (inter + self.epsilon) / (union + self.epsilon)).float()


class F1Loss(nn.Module):
    """model/unet2d/loss.py:46-57."""

    def __init__(self, epsilon=1e-7):
        super().__init__()
        self.epsilon = epsilon

    def forward(self, inputs, targets):
        s = seg_sums(inputs, targets, "sigmoid").sum(dim=0)
        tp = s[S_PT]
        precision = tp / (s[S_P] + self.epsilon)
        recall = tp / (s[S_T] + self.epsilon)
        return (1 - 2 * (precision * recall) / (precision + recall + self.epsilon)).float()


_CONSTS: dict = {}


def _device_const(key, device, dtype, make) -> torch.Tensor:
    """Small constant tensors live on the device once per (key, device): a host -> device copy in every step costs a
    blocking transfer and cannot sit inside a CUDA-graph capture (GraphedTrainStep)."""
    k = (key, device, dtype)
    t = _CONSTS.get(k)
    if t is None:
        t = _CONSTS[k] = torch.tensor(make(), dtype=dtype, device=device)
    return t


class _MSSSIMMeans(torch.autograd.Function):
    """(5, n) per-sample means of the MS-SSIM scale maps of (sigmoid(logits), targets)."""

    @staticmethod
    def forward(ctx, logits: torch.Tensor, targets: torch.Tensor, data_range: float):
        logits = logits.contiguous().float()
        targets = targets.contiguous().float()
        sums, pyr_x, pyr_y = ops.msssim_fwd(logits, targets, data_range)
        ctx.save_for_backward(pyr_x, pyr_y)
        ctx.shape, ctx.data_range = tuple(logits.shape), data_range
        return (sums / _device_const(("counts", logits.shape[2], logits.shape[3]), sums.device, torch.float64,
                                     lambda: ops.msssim_valid_counts(logits.shape[2], logits.shape[3]))[:, None]).float()

    @staticmethod
    def backward(ctx, gmeans):
        pyr_x, pyr_y = ctx.saved_tensors
        return ops.msssim_bwd(ctx.shape, pyr_x, pyr_y, gmeans, ctx.data_range), None, None


MS_SSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


class MSSSIMLoss(nn.Module):
    """model/unet2d/loss.py:21-28: ``1 - MS_SSIM(data_range=1.0, size_average=True, channel=1)(sigmoid(x), t)``
    (pytorch-msssim==1.0.0).  The filtering, the scale pyramid and their adjoints run in
    ``b200_msssim_fwd/bwd``; the five numbers per sample are combined here exactly as the package does
    (ReLU, weights as exponents, product over scales, batch mean)."""

    def __init__(self, data_range: float = 1.0):
        super().__init__()
        self.data_range = data_range

    def forward(self, inputs, targets):
        if inputs.shape != targets.shape:
            raise ValueError("Input images should have the same dimensions.")
        means = _MSSSIMMeans.apply(inputs, targets, self.data_range)          # (5, n)
        weights = _device_const("weights", means.device, means.dtype, lambda: MS_SSIM_WEIGHTS).view(-1, 1)
        # product over the five scales written out: torch.prod's backward inspects its input for zeros on the host
        # (a device -> host sync in every step, and it invalidates a CUDA-graph capture of the step)
        v = torch.relu(means) ** weights
        ms_ssim_val = v[0] * v[1] * v[2] * v[3] * v[4]
        return 1 - ms_ssim_val.mean()


class SegmentationLoss(nn.Module):
    """model/unet2d/loss.py:59-70: ``F1 + MS-SSIM + IoU``.  The F1 and IoU terms come out of ONE pass over
    the logits (the six-sum kernel), the MS-SSIM term from ``MSSSIMLoss``.  ``ms_ssim=False`` drops that
    term (images not larger than 160 pixels, where the reference's package asserts), ``ms_ssim_loss`` replaces it."""

    def __init__(self, ms_ssim_loss: Optional[nn.Module] = None, epsilon: float = 1e-7, ms_ssim: bool = True):
        super().__init__()
        self.f1_loss = F1Loss(epsilon)
        self.iou_loss = IoULoss(epsilon)
        self.ms_ssim_loss = ms_ssim_loss if ms_ssim_loss is not None else (MSSSIMLoss() if ms_ssim else None)
        self.epsilon = epsilon

    def forward(self, inputs, targets):
        s = seg_sums(inputs, targets, "sigmoid").sum(dim=0)
        tp = s[S_PT]
        precision = tp / (s[S_P] + self.epsilon)
        recall = tp / (s[S_T] + self.epsilon)
        f1 = 1 - 2 * (precision * recall) / (precision + recall + self.epsilon)
        union = s[S_P] + s[S_T] - tp
        iou = 1 - (tp + self.epsilon) / (union + self.epsilon)
        loss = (f1 + iou).float()
        if self.ms_ssim_loss is not None:
            loss = loss + self.ms_ssim_loss(inputs, targets)
        return loss


def flatten(tensor: torch.Tensor) -> torch.Tensor:
    """model/unet3d/losses.py:258-270: (N, C, D, H, W) -> (C, N * D * H * W).  Public helper kept for
    drop-in callers; the fused loss kernels reduce in place and never materialise this copy."""
    c = tensor.size(1)
    axis_order = (1, 0) + tuple(range(2, tensor.dim()))
    return tensor.permute(axis_order).contiguous().view(c, -1)


def get_loss_criterion(config):
    """losses.py:273-306 for the losses on the path; anything else is left to stock PyTorch by the
    caller (GeneralizedDice / weighted CE / pixel-wise CE are out of scope)."""
    assert "loss" in config, "Could not find loss function configuration"
    loss_config = dict(config["loss"])
    name = loss_config.pop("name")
    weight = loss_config.pop("weight", None)
    if weight is not None:
        weight = torch.tensor(weight)
    if name == "BCEWithLogitsLoss":
        return BCEWithLogitsLoss()
    if name == "BCEDiceLoss":
        return BCEDiceLoss(alpha=loss_config.get("alpha", 1.0), beta=loss_config.get("beta", 1.0))
    if name == "DiceLoss":
        return DiceLoss(weight=weight, normalization=loss_config.get("normalization", "sigmoid"))
    if name == "CrossEntropyLoss":
        return CrossEntropyLoss()
    raise RuntimeError(f"Unsupported loss function: '{name}'")
