"""Functional fp32 restatement of the reference networks (TEST INFRASTRUCTURE, see oracle/__init__).

Every function takes the reference's own ``state_dict`` (same keys, same tensor layouts) and plain
tensors, so weights trained or initialised by either side can be replayed through the other.
Citations are file:line relative to the reference root.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------- 2D U-Net
def double_convolution(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """model/unet2d/layers.py:103-133 — conv3x3(pad 1, bias) + ReLU, twice, no normalisation."""
    x = F.relu(F.conv2d(x, sd[p + "first.weight"], sd[p + "first.bias"], padding=1))
    return F.relu(F.conv2d(x, sd[p + "second.weight"], sd[p + "second.bias"], padding=1))


def center_crop_to(skip: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """torchvision center_crop as used at layers.py:186 (a no-op for sizes divisible by 16)."""
    sh, sw = skip.shape[-2:]
    if (sh, sw) == (h, w):
        return skip
    top = int(round((sh - h) / 2.0))
    left = int(round((sw - w) / 2.0))
    return skip[..., top:top + h, left:left + w]


def unet2d_forward(sd: SD, images: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """model/unet2d/unet.py:95-128 — 4x(double conv, 2x2 max-pool), bottleneck, 4x(ConvTranspose2d
    k2 s2, cat([up, skip]), double conv), 1x1 head; returns logits."""
    x = images
    skips: List[torch.Tensor] = []
    for i in range(4):
        x = double_convolution(sd, f"{prefix}down_conv.{i}.", x)
        skips.append(x)
        x = F.max_pool2d(x, 2)                                            # layers.py:147
    x = double_convolution(sd, f"{prefix}middle_conv.", x)
    for i in range(4):
        x = F.conv_transpose2d(x, sd[f"{prefix}up_sample.{i}.up.weight"],
                               sd[f"{prefix}up_sample.{i}.up.bias"], stride=2)   # layers.py:165
        skip = center_crop_to(skips.pop(), x.shape[2], x.shape[3])
        x = torch.cat([x, skip], dim=1)                                   # layers.py:188
        x = double_convolution(sd, f"{prefix}up_conv.{i}.", x)
    return F.conv2d(x, sd[prefix + "final_conv.weight"], sd[prefix + "final_conv.bias"])


def unet_conv2(sd: SD, p: str, x: torch.Tensor, n: int = 2, training: bool = True,
               is_batchnorm: bool = True) -> torch.Tensor:
    """model/unet2d/layers.py:8-46 — n x (Conv(bias) -> BatchNorm2d -> ReLU)."""
    for i in range(1, n + 1):
        q = f"{p}conv{i}."
        x = F.conv2d(x, sd[q + "0.weight"], sd[q + "0.bias"], padding=1)
        if is_batchnorm:
            x = F.batch_norm(x, sd[q + "1.running_mean"], sd[q + "1.running_var"], sd[q + "1.weight"],
                             sd[q + "1.bias"], training=training, momentum=0.1, eps=1e-5)
        x = F.relu(x)
    return x


def unet_model_loss(logits: torch.Tensor, labels: torch.Tensor, out_channels: int) -> torch.Tensor:
    """model/unet2d/unet.py:1184-1188,1212 — BCEWithLogits (C == 1) or CrossEntropy (C > 1), mean."""
    if out_channels > 1:
        return F.cross_entropy(logits, labels)
    return F.binary_cross_entropy_with_logits(logits, labels)


def iou_loss(logits: torch.Tensor, targets: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """model/unet2d/loss.py:32-42."""
    p = torch.sigmoid(logits)
    inter = (p * targets).sum()
    union = p.sum() + targets.sum() - inter
    return 1 - (inter + eps) / (union + eps)


def f1_loss(logits: torch.Tensor, targets: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """model/unet2d/loss.py:46-57."""
    p = torch.sigmoid(logits)
    tp = (p * targets).sum()
    precision = tp / (p.sum() + eps)
    recall = tp / (targets.sum() + eps)
    return 1 - 2 * (precision * recall) / (precision + recall + eps)


# ----------------------------------------------------------------------------------------- 3D family
def features_per_level(f_maps, num_levels: int) -> List[int]:
    """model/unet3d/utils.py:109-110."""
    if isinstance(f_maps, int):
        return [f_maps * 2 ** k for k in range(num_levels)]
    return list(f_maps)


def single_conv(sd: SD, p: str, x: torch.Tensor, order: str, num_groups: int, is3d: bool,
                training: bool, dropout_prob: float = 0.1) -> torch.Tensor:
    """model/unet3d/buildingblocks.py:14-159 — the layer_order mini-language over {g, b, c, r, l, e, d, D}
    (dropout letters draw from torch's generator, as the reference's nn.Dropout / nn.Dropout2d do)."""
    conv = F.conv3d if is3d else F.conv2d
    for ch in order:
        if ch == "c":
            x = conv(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"), padding=1)
        elif ch == "g":
            c = x.shape[1]
            g = num_groups if c >= num_groups else 1                       # buildingblocks.py:80-81
            x = F.group_norm(x, g, sd[p + "groupnorm.weight"], sd[p + "groupnorm.bias"], eps=1e-5)
        elif ch == "b":
            x = F.batch_norm(x, sd[p + "batchnorm.running_mean"], sd[p + "batchnorm.running_var"],
                             sd[p + "batchnorm.weight"], sd[p + "batchnorm.bias"], training=training,
                             momentum=0.1, eps=1e-5)
        elif ch == "r":
            x = F.relu(x)
        elif ch == "l":
            x = F.leaky_relu(x)                                            # buildingblocks.py:57 (slope 0.01)
        elif ch == "e":
            x = F.elu(x)                                                   # :59
        elif ch == "d":
            x = F.dropout(x, dropout_prob, training)                       # :104-105
        elif ch == "D":
            x = F.dropout2d(x, dropout_prob, training) if x.dim() == 4 else \
                F.dropout3d(x, dropout_prob, training)                     # :106-107 (channel-wise on 5-D input)
        else:
            raise ValueError(f"layer_order letter {ch!r} is outside the oracle's scope")
    return x


def double_conv(sd: SD, p: str, x, order, num_groups, is3d, training):
    """buildingblocks.py:162-252 (channel rule lives in the weights' shapes)."""
    x = single_conv(sd, p + "SingleConv1.", x, order, num_groups, is3d, training)
    return single_conv(sd, p + "SingleConv2.", x, order, num_groups, is3d, training)


def unet3d_forward(sd: SD, x: torch.Tensor, *, num_levels: int = 4, layer_order: str = "gcr",
                   num_groups: int = 8, upsample: str = "default", is3d: bool = True,
                   training: bool = True, prefix: str = "") -> torch.Tensor:
    """model/unet3d/model.py:125-151 with Encoder (buildingblocks.py:435-439) and Decoder
    (:539-550): pool -> DoubleConv per level; decoder = upsample to the skip's size,
    cat((skip, x)), DoubleConv; 1x1 head; returns logits (final activation is disabled, F8)."""
    pool = F.max_pool3d if is3d else F.max_pool2d
    feats: List[torch.Tensor] = []
    for i in range(num_levels):
        if i > 0:
            x = pool(x, 2)
        x = double_conv(sd, f"{prefix}encoders.{i}.basic_module.", x, layer_order, num_groups, is3d,
                        training)
        feats.insert(0, x)
    for i, skip in enumerate(feats[1:]):
        size = skip.shape[2:]
        if upsample == "deconv":
            w = sd[f"{prefix}decoders.{i}.upsampling.upsample.conv_transposed.weight"]
            ct = F.conv_transpose3d if is3d else F.conv_transpose2d
            x = ct(x, w, None, stride=2, padding=1)                        # buildingblocks.py:710-717
            x = F.interpolate(x, size=size)                                # :703 (nearest)
        else:
            x = F.interpolate(x, size=size, mode="nearest")                # :671-673
        x = torch.cat((skip, x), dim=1)                                    # :547
        x = double_conv(sd, f"{prefix}decoders.{i}.basic_module.", x, layer_order, num_groups, is3d,
                        training)
    conv = F.conv3d if is3d else F.conv2d
    return conv(x, sd[prefix + "final_conv.weight"], sd[prefix + "final_conv.bias"])


def resnet_block(sd: SD, p: str, x, order, num_groups, is3d, training):
    """buildingblocks.py:255-325 ResNetBlock: residual = conv1(x) (1x1 conv iff the channel count changes),
    conv2 (SingleConv ``order``), conv3 (``order`` without its non-linearity), add, ReLU; with ``se_module.*``
    weights present, ResNetBlockSE (:328-362): the scSE gate of model/unet3d/se.py:18-113 on the result."""
    conv = F.conv3d if is3d else F.conv2d
    res = conv(x, sd[p + "conv1.weight"], sd[p + "conv1.bias"]) if p + "conv1.weight" in sd else x
    n_order = order
    for ch in "rel":
        n_order = n_order.replace(ch, "")
    out = single_conv(sd, p + "conv2.", res, order, num_groups, is3d, training)
    out = single_conv(sd, p + "conv3.", out, n_order, num_groups, is3d, training)
    out = out + res
    if "l" in order:                                                # buildingblocks.py:306-311
        out = F.leaky_relu(out, 0.1)
    elif "e" in order:
        out = F.elu(out)
    else:
        out = F.relu(out)
    if any(k.startswith(p + "se_module.") for k in sd):             # ResNetBlockSE, buildingblocks.py:328-362
        out = se_layer(sd, p + "se_module.", out)
    return out


def se_layer(sd: SD, q: str, x: torch.Tensor) -> torch.Tensor:
    """model/unet3d/se.py: ChannelSELayer3D (:41-57, keys ``fc1/fc2``), SpatialSELayer3D (:75-97, key ``conv``) or
    ChannelSpatialSELayer3D (:111-113, keys ``cSE.* / sSE.*``: element-wise max of the two)."""
    def cse(k):
        squeeze = x.mean(dim=(2, 3, 4))                             # :46 AdaptiveAvgPool3d(1)
        a = torch.sigmoid(F.linear(F.relu(F.linear(squeeze, sd[k + "fc1.weight"], sd[k + "fc1.bias"])),
                                   sd[k + "fc2.weight"], sd[k + "fc2.bias"]))          # :49-50
        return x * a[:, :, None, None, None]                        # :52

    def sse(k):
        return x * torch.sigmoid(F.conv3d(x, sd[k + "conv.weight"], sd[k + "conv.bias"]))   # :88-93

    if q + "cSE.fc1.weight" in sd:
        return torch.max(cse(q + "cSE."), sse(q + "sSE."))          # :112
    if q + "fc1.weight" in sd:
        return cse(q)
    return sse(q)


def decoder_forward(sd: SD, p: str, skip: torch.Tensor, x: torch.Tensor, *, order: str = "gcr",
                    num_groups: int = 8, is3d: bool = True, training: bool = True,
                    residual: bool = False) -> torch.Tensor:
    """buildingblocks.py:539-550 ``Decoder.forward(encoder_features, x)`` stand-alone: DoubleConv decoders
    up-sample (nearest, or ConvTranspose k3 s2 p1 + nearest resize when a ``conv_transposed`` weight exists)
    and concatenate ``(skip, x)``; ResNetBlock decoders (``residual``) up-sample by the transposed convolution
    and sum."""
    key = p + "upsampling.upsample.conv_transposed.weight"
    if key in sd:
        ct = F.conv_transpose3d if is3d else F.conv_transpose2d
        x = F.interpolate(ct(x, sd[key], None, stride=2, padding=1), size=skip.shape[2:])
    else:
        x = F.interpolate(x, size=skip.shape[2:], mode="nearest")
    if residual:
        return resnet_block(sd, p + "basic_module.", skip + x, order, num_groups, is3d, training)
    return double_conv(sd, p + "basic_module.", torch.cat((skip, x), dim=1), order, num_groups, is3d, training)


def encoder_forward(sd: SD, p: str, x: torch.Tensor, *, order: str = "gcr", num_groups: int = 8,
                    is3d: bool = True, training: bool = True, pool: Optional[str] = "max",
                    residual: bool = False) -> torch.Tensor:
    """buildingblocks.py:435-439 ``Encoder.forward``: optional 2x max / avg pooling (:409-418), basic module."""
    if pool == "max":
        x = (F.max_pool3d if is3d else F.max_pool2d)(x, 2)
    elif pool == "avg":
        x = (F.avg_pool3d if is3d else F.avg_pool2d)(x, 2)
    if residual:
        return resnet_block(sd, p + "basic_module.", x, order, num_groups, is3d, training)
    return double_conv(sd, p + "basic_module.", x, order, num_groups, is3d, training)


def residual_unet3d_forward(sd: SD, x: torch.Tensor, *, num_levels: int = 5, layer_order: str = "gcr",
                            num_groups: int = 8, is3d: bool = True, training: bool = True) -> torch.Tensor:
    """model/unet3d/model.py:197-238 ResidualUNet3D (and, with ``is3d=False``, ResidualUNet2D :323-359):
    ResNetBlock levels, decoder = ConvTranspose(k3, s2, p1, no bias) + nearest resize to the skip's size
    (buildingblocks.py:703-717), summation joining (:549), ResNetBlock(out, out); 1x1 head; logits."""
    pool = F.max_pool3d if is3d else F.max_pool2d
    feats: List[torch.Tensor] = []
    for i in range(num_levels):
        if i > 0:
            x = pool(x, 2)
        x = resnet_block(sd, f"encoders.{i}.basic_module.", x, layer_order, num_groups, is3d, training)
        feats.insert(0, x)
    ct = F.conv_transpose3d if is3d else F.conv_transpose2d
    for i, skip in enumerate(feats[1:]):
        w = sd[f"decoders.{i}.upsampling.upsample.conv_transposed.weight"]
        x = F.interpolate(ct(x, w, None, stride=2, padding=1), size=skip.shape[2:])
        x = skip + x
        x = resnet_block(sd, f"decoders.{i}.basic_module.", x, layer_order, num_groups, is3d, training)
    conv = F.conv3d if is3d else F.conv2d
    return conv(x, sd["final_conv.weight"], sd["final_conv.bias"])


# ----------------------------------------------------------------------------------------- 3D losses
def flatten_channels_first(t: torch.Tensor) -> torch.Tensor:
    """model/unet3d/losses.py:258-270 — (N, C, ...) -> (C, N * ...)."""
    c = t.shape[1]
    return t.transpose(0, 1).reshape(c, -1)


def per_channel_dice(prob: torch.Tensor, target: torch.Tensor, eps: float = 1e-6,
                     weight: Optional[torch.Tensor] = None) -> torch.Tensor:
    """losses.py:7-33 — 2*sum(p t) / clamp(sum p^2 + sum t^2, eps) per channel."""
    assert prob.shape == target.shape
    p = flatten_channels_first(prob)
    t = flatten_channels_first(target).float()
    inter = (p * t).sum(-1)
    if weight is not None:
        inter = weight * inter
    denom = (p * p).sum(-1) + (t * t).sum(-1)
    return 2 * (inter / denom.clamp(min=eps))


def dice_loss(logits, target, normalization: str = "sigmoid", weight=None) -> torch.Tensor:
    """losses.py:83-129."""
    if normalization == "sigmoid":
        p = torch.sigmoid(logits)
    elif normalization == "softmax":
        p = torch.softmax(logits, dim=1)
    else:
        p = logits
    return 1.0 - per_channel_dice(p, target, weight=weight).mean()


def bce_dice_loss(logits, target, alpha: float = 1.0, beta: float = 1.0) -> torch.Tensor:
    """losses.py:167-178."""
    return alpha * F.binary_cross_entropy_with_logits(logits, target) + beta * dice_loss(logits, target)


def dice_coefficient(prob, target, eps: float = 1e-6) -> torch.Tensor:
    """model/unet3d/metrics.py:15-29."""
    return per_channel_dice(prob, target, eps=eps).mean()


def mean_iou(prob: torch.Tensor, target: torch.Tensor, skip_channels: Sequence[int] = ()) -> torch.Tensor:
    """model/unet3d/metrics.py:32-104 (ignore_index = None, 5-D one-hot target)."""
    assert prob.dim() == 5 and prob.shape == target.shape
    n_classes = prob.shape[1]
    per_batch = []
    for _p, _t in zip(prob, target):
        if n_classes == 1:
            pred = (_p > 0.5).to(torch.uint8)
        else:
            pred = torch.zeros_like(_p, dtype=torch.uint8).scatter_(0, _p.argmax(0, keepdim=True), 1)
        tb = _t.to(torch.uint8)
        ious = []
        for c in range(n_classes):
            if c in skip_channels:
                continue
            inter = (pred[c] & tb[c]).sum().float()
            union = (pred[c] | tb[c]).sum().float().clamp(min=1e-8)
            ious.append(inter / union)
        per_batch.append(torch.stack(ious).mean())
    return torch.stack(per_batch).mean()
