"""MS-SSIM as the reference's SegmentationLoss uses it (model/unet2d/loss.py:21-28) — TEST INFRASTRUCTURE
(see oracle/__init__).

PARITY UNPINNED for this term: the arithmetic lives in the third-party package ``pytorch-msssim==1.0.0``
(requirements.txt:127), which is neither vendored under /root/reference nor installed in this image, so the
functions below restate the package's published algorithm (Wang et al. 2003 multi-scale SSIM as implemented by
VainF/pytorch-msssim: ``_fspecial_gauss_1d``, ``gaussian_filter``, ``_ssim``, ``ms_ssim``) and are anchored on
the reference's call site — ``MS_SSIM(data_range=1.0, size_average=True, channel=1)`` applied to
``sigmoid(logits)`` and the targets, loss = 1 - value — and on analytic known answers (identical images -> 1,
constant images, the package's size assertion), not on outputs of the package itself.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def fspecial_gauss_1d(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    """pytorch_msssim._fspecial_gauss_1d: normalised 1-D Gaussian, shape (1, 1, size)."""
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g = g / g.sum()
    return g.unsqueeze(0).unsqueeze(0)


def gaussian_filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    """pytorch_msssim.gaussian_filter: separable VALID (no padding) depth-wise blur of (N, C, H, W)."""
    c = x.shape[1]
    w = win.to(x.dtype).repeat(c, 1, 1).unsqueeze(2)           # (C, 1, 1, size)
    out = F.conv2d(x, w.transpose(2, 3), groups=c)              # along H
    return F.conv2d(out, w, groups=c)                           # along W


def ssim_maps(x, y, data_range=1.0, win=None, k=(0.01, 0.03)):
    """pytorch_msssim._ssim: per-(sample, channel) means of the ssim map and of the contrast-structure map."""
    win = fspecial_gauss_1d() if win is None else win
    c1, c2 = (k[0] * data_range) ** 2, (k[1] * data_range) ** 2
    mu1, mu2 = gaussian_filter(x, win), gaussian_filter(y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = gaussian_filter(x * x, win) - mu1_sq
    sigma2_sq = gaussian_filter(y * y, win) - mu2_sq
    sigma12 = gaussian_filter(x * y, win) - mu1_mu2
    cs_map = (2 * sigma12 + c2) / (sigma1_sq + sigma2_sq + c2)
    ssim_map = ((2 * mu1_mu2 + c1) / (mu1_sq + mu2_sq + c1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def ms_ssim(x, y, data_range=1.0, size_average=True, weights=WEIGHTS):
    """pytorch_msssim.ms_ssim: 5 scales, cs of scales 0-3 and ssim of scale 4, each ReLU'd and raised to its
    weight, product over scales; avg_pool2d(kernel 2, padding = size % 2) between scales."""
    if x.shape != y.shape:
        raise ValueError("Input images should have the same dimensions.")
    assert min(x.shape[-2:]) > (11 - 1) * 2 ** 4, "Image size should be larger than 160 due to the 4 downsamplings"
    w = torch.tensor(weights, dtype=x.dtype)
    levels = w.shape[0]
    mcs = []
    for i in range(levels):
        ssim_pc, cs = ssim_maps(x, y, data_range)
        if i < levels - 1:
            mcs.append(torch.relu(cs))
            padding = [s % 2 for s in x.shape[2:]]
            x = F.avg_pool2d(x, kernel_size=2, padding=padding)
            y = F.avg_pool2d(y, kernel_size=2, padding=padding)
    vals = torch.stack(mcs + [torch.relu(ssim_pc)], dim=0)      # (level, batch, channel)
    ms = torch.prod(vals ** w.view(-1, 1, 1), dim=0)
    return ms.mean() if size_average else ms.mean(1)


def msssim_loss(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """loss.py:21-28 MSSSIMLoss."""
    return 1 - ms_ssim(torch.sigmoid(logits), targets, data_range=1.0, size_average=True)


def segmentation_loss(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """loss.py:59-70 SegmentationLoss = F1 + MS-SSIM + IoU."""
    from .unet import f1_loss, iou_loss
    return f1_loss(logits, targets) + msssim_loss(logits, targets) + iou_loss(logits, targets)
