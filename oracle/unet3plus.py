"""Functional fp32 restatement of UNet 3+ (model/unet2d/unet.py:136-446) — TEST INFRASTRUCTURE (see
oracle/__init__).  Takes the reference's own ``state_dict`` (same keys) and plain tensors."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .unet import unet_conv2

SD = Dict[str, torch.Tensor]


def branch_name(level: int, stage: int) -> str:
    """unet.py:171-315: h{l}_PT_hd{s} (pooled encoder map), h{l}_Cat_hd{s} (same scale), hd{l}_UT_hd{s}
    (bilinearly up-sampled deeper map)."""
    if level < stage:
        return f"h{level}_PT_hd{stage}"
    if level == stage:
        return f"h{level}_Cat_hd{stage}"
    return f"hd{level}_UT_hd{stage}"


def _bf(t: torch.Tensor) -> torch.Tensor:
    """Round to bf16 with a straight-through gradient (storage-precision emulation)."""
    return t + (t.bfloat16().float() - t).detach()


ALL_SITES = frozenset({"w", "x", "pre_bn", "act", "up", "head"})


def _sites(emulate_bf16):
    """``emulate_bf16``: False, True (every storage site) or a set of site names — "w" filter packs, "x" conv inputs,
    "pre_bn" conv outputs in front of a BatchNorm, "act" normalised + ReLU outputs, "up" bilinear outputs, "head" the
    output convs — for the per-site error budget (tools/u3p_error_budget.py)."""
    if emulate_bf16 is True:
        return ALL_SITES
    return frozenset(emulate_bf16) if emulate_bf16 else frozenset()


def conv_bn_relu(sd: SD, conv: str, bn: str, x: torch.Tensor, training: bool, emulate_bf16=False):
    sites = _sites(emulate_bf16)
    w = sd[conv + ".weight"]
    if "x" in sites:
        x = _bf(x)
    if "w" in sites:
        w = _bf(w)
    x = F.conv2d(x, w, sd[conv + ".bias"], padding=1)
    if "pre_bn" in sites:
        x = _bf(x)
    x = F.batch_norm(x, sd[bn + ".running_mean"], sd[bn + ".running_var"], sd[bn + ".weight"], sd[bn + ".bias"],
                     training=training, momentum=0.1, eps=1e-5)
    x = F.relu(x)
    return _bf(x) if "act" in sites else x


def unet3plus_forward(sd: SD, x: torch.Tensor, training: bool = True, emulate_bf16=False,
                      deep_sup: bool = False, cgm: bool = False):
    """unet.py:335-446: encoder (unetConv2 + MaxPool2d(2)) -> four full-scale decoder stages -> 3x3 output
    conv; returns logits (the sigmoid at :445 is commented out).  ``emulate_bf16`` rounds weights and every
    stored activation to bf16 where the B200 path stores bf16 (fp32 accumulation kept): the yardstick that
    separates storage-precision noise — large for random weights on a 19-layer BatchNorm net — from bugs.
    ``deep_sup``: UNet_3Plus_DeepSup (unet.py:454-787) — the tuple (d1..d5) of per-scale logits, each
    bilinearly up-sampled to the input size.  ``cgm`` (with deep_sup): UNet_3Plus_DeepSup_CGM (unet.py:795-1153)
    — the five maps gated per sample by the classification branch, as probabilities."""
    sites = _sites(emulate_bf16)
    h = {}
    for lvl in range(1, 6):
        if lvl > 1:
            x = F.max_pool2d(x, 2)                                             # unet.py:343-353
        if sites:
            for i in (1, 2):
                x = conv_bn_relu(sd, f"conv{lvl}.conv{i}.0", f"conv{lvl}.conv{i}.1", x, training, sites)
        else:
            x = unet_conv2(sd, f"conv{lvl}.", x, training=training)
        h[lvl] = x
    hd = {5: h[5]}
    for stage in (4, 3, 2, 1):
        parts = []
        for level in range(1, 6):
            name = branch_name(level, stage)
            if level < stage:
                k = 2 ** (stage - level)
                src = F.max_pool2d(h[level], k, k, ceil_mode=True)             # unet.py:175-187
            elif level == stage:
                src = h[level]
            else:                                                              # nn.Upsample(bilinear), :198
                src = F.interpolate(hd[level], scale_factor=2 ** (level - stage), mode="bilinear")
                if "up" in sites:
                    src = _bf(src)
            parts.append(conv_bn_relu(sd, name + "_conv", name + "_bn", src, training, sites))
        hd[stage] = conv_bn_relu(sd, f"conv{stage}d_1", f"bn{stage}d_1", torch.cat(parts, 1), training, sites)

    def head(i):
        w = sd[f"outconv{i}.weight"]
        y = F.conv2d(hd[i], _bf(w) if "head" in sites else w, sd[f"outconv{i}.bias"], padding=1)
        return _bf(y) if "head" in sites else y

    if not deep_sup:
        return head(1)
    outs = [head(1)]                                                           # unet.py:760-787
    for i in range(2, 6):
        y = F.interpolate(head(i), scale_factor=2 ** (i - 1), mode="bilinear")
        outs.append(_bf(y) if "head" in sites else y)
    if not cgm:
        return tuple(outs)
    # class-guided module, unet.py:999-1004, 1035-1038, 1012-1017, 1147-1153 (dropout: identity, i.e. eval mode
    # or p = 0; its random mask cannot be pinned)
    cls = torch.sigmoid(F.conv2d(hd[5], sd["cls.1.weight"], sd["cls.1.bias"]).amax(dim=(2, 3)))     # (B, 2)
    gate = cls.argmax(dim=1).float().view(-1, 1, 1, 1)
    return tuple(torch.sigmoid(d * gate) for d in outs)


def fill_state_dict_(sd: SD, seed: int) -> None:
    """Deterministic, init-independent weights for parity runs: the reference model, the oracle and the
    product module get the SAME values for the same keys regardless of how each was initialised (the 27 M
    parameters of UNet 3+ are never stored in a fixture).  He-style scale for conv weights, BN gamma near 1."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        v = sd[k]
        if not v.is_floating_point():
            v.zero_()
        elif k.endswith("running_mean"):
            v.zero_()
        elif k.endswith("running_var"):
            v.fill_(1.0)
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5)
        elif k.endswith("_bn.weight") or k.endswith(".1.weight") or (k.startswith("bn") and k.endswith(".weight")):
            v.copy_(1.0 + 0.1 * torch.randn(v.shape, generator=g))
        else:
            v.copy_(0.05 * torch.randn(v.shape, generator=g))
