"""Drop-in 3D (and 2D-by-the-same-factory) U-Net classes of the reference's pytorch-3dunet family:
same class names, constructor arguments, child names and ``state_dict`` keys as
model/unet3d/buildingblocks.py and model/unet3d/model.py, executed by the sm_100a kernels.

In scope: every ``layer_order`` letter {c, g, b, r, l, e, d, D}, max- and
avg-pool encoders, nearest or 'deconv' (ConvTranspose k3 s2 p1 + nearest resize) decoders with concat or
summation joining, DoubleConv / ResNetBlock / ResNetBlockSE (scse, cse, sse) levels, 1x1 head; every block
class also runs stand-alone (``forward``) through the same kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import partial
from typing import Optional

import torch
from torch import nn
from torch.nn import functional as F

from . import engine
from .losses import get_loss_criterion

try:
    from transformers import PretrainedConfig, PreTrainedModel
    from transformers.utils import ModelOutput
except Exception:  # pragma: no cover
    PretrainedConfig = PreTrainedModel = ModelOutput = None


def number_of_features_per_level(init_channel_number, num_levels):
    """model/unet3d/utils.py:109-110."""
    return [init_channel_number * 2 ** k for k in range(num_levels)]


def create_conv(in_channels, out_channels, kernel_size, order, num_groups, padding, dropout_prob, is3d):
    """buildingblocks.py:14-113 — named (module) list for one layer; the conv has a bias only when no
    norm letter is present, GroupNorm falls back to one group for narrow inputs."""
    assert "c" in order, "Conv layer MUST be present"
    assert order[0] not in "rle", "Non-linearity cannot be the first operation in the layer"
    if kernel_size != 3 or padding != 1:
        raise NotImplementedError("the B200 path implements kernel_size=3, padding=1 convolutions")
    Conv = nn.Conv3d if is3d else nn.Conv2d
    BN = nn.BatchNorm3d if is3d else nn.BatchNorm2d
    has_norm = "g" in order or "b" in order
    before_conv = True
    mods = []
    for ch in order:
        if ch == "c":
            mods.append(("conv", Conv(in_channels, out_channels, kernel_size, padding=padding,
                                      bias=not has_norm)))
            before_conv = False
        elif ch == "r":
            mods.append(("ReLU", nn.ReLU(inplace=True)))
        elif ch == "g":
            nch = in_channels if before_conv else out_channels
            groups = num_groups if nch >= num_groups else 1
            assert nch % groups == 0, (
                f"Expected number of channels in input to be divisible by num_groups. "
                f"num_channels={nch}, num_groups={groups}")
            mods.append(("groupnorm", nn.GroupNorm(num_groups=groups, num_channels=nch)))
        elif ch == "b":
            mods.append(("batchnorm", BN(in_channels if before_conv else out_channels)))
        elif ch == "l":
            mods.append(("LeakyReLU", nn.LeakyReLU(inplace=True)))
        elif ch == "e":
            mods.append(("ELU", nn.ELU(inplace=True)))
        elif ch == "d":
            mods.append(("dropout", nn.Dropout(p=dropout_prob)))
        elif ch == "D":
            mods.append(("dropout2d", nn.Dropout2d(p=dropout_prob)))
        else:
            raise ValueError(f"Unsupported layer type '{ch}'. MUST be one of "
                             "['b', 'g', 'r', 'l', 'e', 'c', 'd', 'D']")
    return mods


class SingleConv(nn.Sequential):
    """buildingblocks.py:116-159."""

    def __init__(self, in_channels, out_channels, kernel_size=3, order="gcr", num_groups=8,
                 padding=1, dropout_prob=0.1, is3d=True):
        super().__init__()
        self.order = order
        for name, module in create_conv(in_channels, out_channels, kernel_size, order, num_groups,
                                        padding, dropout_prob, is3d):
            self.add_module(name, module)

    def forward(self, x):
        return engine.run_block(self, [engine.single_conv_spec(self, self.order)], x)


class DoubleConv(nn.Sequential):
    """buildingblocks.py:162-252 — encoder: in -> max(out/2, in) -> out; decoder: in -> out -> out."""

    def __init__(self, in_channels, out_channels, encoder, kernel_size=3, order="gcr", num_groups=8,
                 padding=1, upscale=2, dropout_prob=0.1, is3d=True):
        super().__init__()
        self.order = order
        if encoder:
            mid = out_channels if upscale == 1 else out_channels // 2
            mid = max(mid, in_channels)
            plan = [(in_channels, mid), (mid, out_channels)]
        else:
            plan = [(in_channels, out_channels), (out_channels, out_channels)]
        probs = dropout_prob if isinstance(dropout_prob, (list, tuple)) else (dropout_prob,) * 2
        for idx, (ci, co) in enumerate(plan):
            self.add_module(f"SingleConv{idx + 1}",
                            SingleConv(ci, co, kernel_size, order, num_groups, padding=padding,
                                       dropout_prob=probs[idx], is3d=is3d))

    def forward(self, x):
        return engine.run_block(self, engine.double_conv_specs(self, self.order), x)


class ResNetBlock(nn.Module):
    """buildingblocks.py:255-325 — ``act(conv3(conv2(r)) + r)`` with ``r = conv1(x)`` (1x1x1 conv when the channel
    count changes); ``act`` = LeakyReLU(0.1) for 'l' orders, ELU for 'e' orders (the class default ``cge``), else
    ReLU.  Inside ``ResidualUNet3D`` / ``ResidualUNet2D`` the whole-network executor (resnet_exec.py) runs it;
    ``forward`` runs the same kernels stand-alone."""

    def __init__(self, in_channels, out_channels, kernel_size=3, order="cge", num_groups=8, is3d=True, **kwargs):
        super().__init__()
        if in_channels != out_channels:
            self.conv1 = (nn.Conv3d if is3d else nn.Conv2d)(in_channels, out_channels, 1)
        else:
            self.conv1 = nn.Identity()
        self.conv2 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=order,
                                num_groups=num_groups, is3d=is3d)
        n_order = order
        for c in "rel":
            n_order = n_order.replace(c, "")
        self.conv3 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=n_order,
                                num_groups=num_groups, is3d=is3d)
        if "l" in order:
            self.non_linearity = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        elif "e" in order:
            self.non_linearity = nn.ELU(inplace=True)
        else:
            self.non_linearity = nn.ReLU(inplace=True)
        self.order = order
        self.is3d = is3d

    def forward(self, x):
        from . import resnet_exec
        return resnet_exec.run_block_standalone(self, x)


class ChannelSELayer3D(nn.Module):
    """se.py:18-57 (parameter container: ``fc1``, ``fc2``) — ``x * sigmoid(fc2(relu(fc1(mean(x)))))``."""

    def __init__(self, num_channels, reduction_ratio=2):
        super().__init__()
        self.reduction_ratio = reduction_ratio
        self.fc1 = nn.Linear(num_channels, num_channels // reduction_ratio, bias=True)
        self.fc2 = nn.Linear(num_channels // reduction_ratio, num_channels, bias=True)

    def forward(self, x):
        from . import resnet_exec
        return resnet_exec.run_se_standalone(self, x)


class SpatialSELayer3D(nn.Module):
    """se.py:60-97 (parameter container: ``conv`` = Conv3d(C, 1, 1)) — ``x * sigmoid(conv(x))``."""

    def __init__(self, num_channels):
        super().__init__()
        self.conv = nn.Conv3d(num_channels, 1, 1)

    def forward(self, x, weights=None):
        if weights is not None:
            raise NotImplementedError("few-shot 'weights' of SpatialSELayer3D (se.py:81-84) are outside the path")
        from . import resnet_exec
        return resnet_exec.run_se_standalone(self, x)


class ChannelSpatialSELayer3D(nn.Module):
    """se.py:100-113: ``max(cSE(x), sSE(x))`` — executed by ``b200_se_*`` inside the residual executor."""

    def __init__(self, num_channels, reduction_ratio=2):
        super().__init__()
        self.cSE = ChannelSELayer3D(num_channels, reduction_ratio)
        self.sSE = SpatialSELayer3D(num_channels)

    def forward(self, x):
        from . import resnet_exec
        return resnet_exec.run_se_standalone(self, x)


class ResNetBlockSE(ResNetBlock):
    """buildingblocks.py:328-362 — ResNetBlock followed by a squeeze-and-excitation gate: ``scse`` (concurrent
    spatial and channel, the variant ``ResidualUNetSE3D`` builds), ``cse`` or ``sse``.  The fused gate kernels use
    ``max(y a, y b) = y max(a, b)``, which needs a non-negative block output: the block's non-linearity must be
    ReLU (orders with 'r', e.g. the networks' ``gcr``); 'l' / 'e' orders raise."""

    def __init__(self, in_channels, out_channels, kernel_size=3, order="cge", num_groups=8, se_module="scse",
                 **kwargs):
        super().__init__(in_channels, out_channels, kernel_size=kernel_size, order=order,
                         num_groups=num_groups, **kwargs)
        assert se_module in ["scse", "cse", "sse"]
        if "l" in order or "e" in order:
            raise NotImplementedError("ResNetBlockSE on the B200 path needs a ReLU block output (order with 'r'): "
                                      "the fused SE gate relies on y >= 0")
        if out_channels > 1024:
            raise NotImplementedError("SE blocks wider than 1024 channels are outside the B200 path")
        if se_module == "scse":
            self.se_module = ChannelSpatialSELayer3D(num_channels=out_channels, reduction_ratio=1)
        elif se_module == "cse":
            self.se_module = ChannelSELayer3D(num_channels=out_channels, reduction_ratio=1)
        else:
            self.se_module = SpatialSELayer3D(num_channels=out_channels)


class Encoder(nn.Module):
    """buildingblocks.py:365-439 — optional 2x max- or avg-pool, then the basic module."""

    def __init__(self, in_channels, out_channels, conv_kernel_size=3, apply_pooling=True,
                 pool_kernel_size=2, pool_type="max", basic_module=DoubleConv, conv_layer_order="gcr",
                 num_groups=8, padding=1, upscale=2, dropout_prob=0.1, is3d=True):
        super().__init__()
        assert pool_type in ["max", "avg"]
        if basic_module not in (DoubleConv, ResNetBlock, ResNetBlockSE):
            raise NotImplementedError("only DoubleConv / ResNetBlock(SE) encoders are on the B200 hot path")
        self.is3d, self.pool_type = is3d, pool_type
        if apply_pooling:
            if pool_kernel_size != 2:
                raise NotImplementedError("the B200 path implements 2x pooling (pool_kernel_size=2)")
            if pool_type == "max":
                self.pooling = nn.MaxPool3d(kernel_size=2) if is3d else nn.MaxPool2d(kernel_size=2)
            else:
                self.pooling = nn.AvgPool3d(kernel_size=2) if is3d else nn.AvgPool2d(kernel_size=2)
        else:
            self.pooling = None
        self.basic_module = basic_module(in_channels, out_channels, encoder=True,
                                         kernel_size=conv_kernel_size, order=conv_layer_order,
                                         num_groups=num_groups, padding=padding, upscale=upscale,
                                         dropout_prob=dropout_prob, is3d=is3d)

    def forward(self, x):  # stand-alone use; inside a network the executor fuses the whole level
        if self.pooling is not None:
            x = engine.run_pool(x, self.is3d, self.pool_type == "avg")   # the nn.*Pool child is only a marker
        return self.basic_module(x)


class AbstractUpsampling(nn.Module):
    """buildingblocks.py:642-656."""

    def __init__(self, upsample):
        super().__init__()
        self.upsample = upsample

    def forward(self, encoder_features, x):
        return self.upsample(x, encoder_features.size()[2:])


class InterpolateUpsampling(AbstractUpsampling):
    """buildingblocks.py:659-673."""

    def __init__(self, mode="nearest"):
        if mode != "nearest":
            raise NotImplementedError("the B200 path implements nearest-neighbour up-sampling")
        super().__init__(partial(self._interpolate, mode=mode))

    @staticmethod
    def _interpolate(x, size, mode):
        return F.interpolate(x, size=size, mode=mode)


class TransposeConvUpsampling(AbstractUpsampling):
    """buildingblocks.py:676-728 — ConvTranspose(k3, s2, p1, no bias) then nearest resize."""

    class Upsample(nn.Module):
        def __init__(self, conv_transposed, is3d):
            super().__init__()
            self.conv_transposed = conv_transposed
            self.is3d = is3d

        def forward(self, x, size):
            return F.interpolate(self.conv_transposed(x), size=size)

    def __init__(self, in_channels, out_channels, kernel_size=3, scale_factor=2, is3d=True):
        if kernel_size != 3 or scale_factor != 2:
            raise NotImplementedError("the B200 path implements kernel 3 / stride 2 deconvolution")
        CT = nn.ConvTranspose3d if is3d else nn.ConvTranspose2d
        ct = CT(in_channels, out_channels, kernel_size=kernel_size, stride=scale_factor, padding=1,
                bias=False)
        super().__init__(self.Upsample(ct, is3d))


class NoUpsampling(AbstractUpsampling):
    def __init__(self):
        super().__init__(self._no_upsampling)

    @staticmethod
    def _no_upsampling(x, size):
        return x


class Decoder(nn.Module):
    """buildingblocks.py:442-550 — upsample to the skip's size, cat((skip, x)), basic module."""

    def __init__(self, in_channels, out_channels, conv_kernel_size=3, scale_factor=2,
                 basic_module=DoubleConv, conv_layer_order="gcr", num_groups=8, padding=1,
                 upsample="default", dropout_prob=0.1, is3d=True):
        super().__init__()
        if basic_module not in (DoubleConv, ResNetBlock, ResNetBlockSE):
            raise NotImplementedError("only DoubleConv / ResNetBlock(SE) decoders are on the B200 hot path")
        if upsample is None or upsample == "none":
            raise NotImplementedError("decoders without up-sampling are outside the B200 hot path")
        concat = True
        if upsample == "default":          # buildingblocks.py:486-501
            if basic_module is DoubleConv:
                upsample = "nearest"
            else:
                upsample, concat = "deconv", False      # learned up-sampling, summation joining,
        if basic_module is not DoubleConv and (upsample != "deconv" or concat):
            raise NotImplementedError("ResNetBlock decoders run with upsample='default' (deconv + summation)")
        if upsample == "deconv":
            self.upsampling = TransposeConvUpsampling(in_channels=in_channels, out_channels=out_channels,
                                                      kernel_size=conv_kernel_size,
                                                      scale_factor=scale_factor, is3d=is3d)
        else:
            self.upsampling = InterpolateUpsampling(mode=upsample)
        self.joining = partial(self._joining, concat=concat)
        if not concat:
            in_channels = out_channels                  # adapt_channels, buildingblocks.py:524-526
        self.basic_module = basic_module(in_channels, out_channels, encoder=False,
                                         kernel_size=conv_kernel_size, order=conv_layer_order,
                                         num_groups=num_groups, padding=padding,
                                         dropout_prob=dropout_prob, is3d=is3d)

        self.concat, self.is3d = concat, is3d

    def forward(self, encoder_features, x):
        """buildingblocks.py:539-543, stand-alone: up-sample ``x`` to the skip's size, join, basic module."""
        from . import resnet_exec
        return resnet_exec.run_decoder_standalone(self, encoder_features, x)

    @staticmethod
    def _joining(encoder_features, x, concat):
        return torch.cat((encoder_features, x), dim=1) if concat else encoder_features + x


def create_encoders(in_channels, f_maps, basic_module, conv_kernel_size, conv_padding, conv_upscale,
                    dropout_prob, layer_order, num_groups, pool_kernel_size, is3d, pool_type="max"):
    """buildingblocks.py:553-601."""
    encoders = []
    for i, out_feature_num in enumerate(f_maps):
        encoders.append(Encoder(in_channels if i == 0 else f_maps[i - 1], out_feature_num,
                                apply_pooling=i > 0, basic_module=basic_module,
                                conv_layer_order=layer_order, conv_kernel_size=conv_kernel_size,
                                num_groups=num_groups, pool_kernel_size=pool_kernel_size, pool_type=pool_type,
                                padding=conv_padding, upscale=conv_upscale, dropout_prob=dropout_prob,
                                is3d=is3d))
    return nn.ModuleList(encoders)


def create_decoders(f_maps, basic_module, conv_kernel_size, conv_padding, layer_order, num_groups,
                    upsample, dropout_prob, is3d):
    """buildingblocks.py:604-639 — in-channels = f[i] + f[i+1] unless the up-sampling is learned."""
    decoders = []
    rev = list(reversed(f_maps))
    for i in range(len(rev) - 1):
        cin = rev[i] + rev[i + 1] if (basic_module == DoubleConv and upsample != "deconv") else rev[i]
        decoders.append(Decoder(cin, rev[i + 1], basic_module=basic_module,
                                conv_layer_order=layer_order, conv_kernel_size=conv_kernel_size,
                                num_groups=num_groups, padding=conv_padding, upsample=upsample,
                                dropout_prob=dropout_prob, is3d=is3d))
    return nn.ModuleList(decoders)


class AbstractUNet(nn.Module):
    """model/unet3d/model.py:13-151. ``forward`` returns logits (the final activation is disabled in
    the reference, model.py:146-149)."""

    def __init__(self, in_channels, out_channels, final_sigmoid, basic_module, f_maps=64,
                 layer_order="gcr", num_groups=8, num_levels=4, is_segmentation=True,
                 conv_kernel_size=3, pool_kernel_size=2, conv_padding=1, conv_upscale=2,
                 upsample="default", dropout_prob=0.1, is3d=True):
        super().__init__()
        if isinstance(f_maps, int):
            f_maps = number_of_features_per_level(f_maps, num_levels=num_levels)
        assert isinstance(f_maps, (list, tuple))
        assert len(f_maps) > 1, "Required at least 2 levels in the U-Net"
        if "g" in layer_order:
            assert num_groups is not None, "num_groups must be specified if GroupNorm is used"
        self.is3d, self.layer_order = is3d, layer_order
        self.encoders = create_encoders(in_channels, f_maps, basic_module, conv_kernel_size,
                                        conv_padding, conv_upscale, dropout_prob, layer_order,
                                        num_groups, pool_kernel_size, is3d)
        self.decoders = create_decoders(f_maps, basic_module, conv_kernel_size, conv_padding,
                                        layer_order, num_groups, upsample, dropout_prob, is3d)
        self.final_conv = (nn.Conv3d if is3d else nn.Conv2d)(f_maps[0], out_channels, 1)
        if is_segmentation:
            self.final_activation = nn.Sigmoid() if final_sigmoid else nn.Softmax(dim=1)
        else:
            self.final_activation = None

        self._kind = "unet3d" if basic_module is DoubleConv else "resunet3d"

    def forward(self, x):
        return engine.run_network(self, self._kind, x)


class ResidualUNet3D(AbstractUNet):
    """model.py:197-238 — ResNetBlock levels, transposed-conv up-sampling, summation joining; 5 levels by
    default.  ``state_dict`` keys as the reference: ``encoders.i.basic_module.{conv1,conv2.*,conv3.*}``,
    ``decoders.i.upsampling.upsample.conv_transposed.weight``, ``final_conv``."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr",
                 num_groups=8, num_levels=5, is_segmentation=True, conv_padding=1, conv_upscale=2,
                 upsample="default", dropout_prob=0.1, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels,
                         final_sigmoid=final_sigmoid, basic_module=ResNetBlock, f_maps=f_maps,
                         layer_order=layer_order, num_groups=num_groups, num_levels=num_levels,
                         is_segmentation=is_segmentation, conv_padding=conv_padding,
                         conv_upscale=conv_upscale, upsample=upsample, dropout_prob=dropout_prob, is3d=True)


class ResidualUNetSE3D(AbstractUNet):
    """model.py:241-280 — ResidualUNet3D with ResNetBlockSE levels (scSE gate after every block)."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr",
                 num_groups=8, num_levels=5, is_segmentation=True, conv_padding=1, conv_upscale=2,
                 upsample="default", dropout_prob=0.1, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels,
                         final_sigmoid=final_sigmoid, basic_module=ResNetBlockSE, f_maps=f_maps,
                         layer_order=layer_order, num_groups=num_groups, num_levels=num_levels,
                         is_segmentation=is_segmentation, conv_padding=conv_padding,
                         conv_upscale=conv_upscale, upsample=upsample, dropout_prob=dropout_prob, is3d=True)


class ResidualUNet2D(AbstractUNet):
    """model.py:323-359 — the residual U-Net with 2D convolutions / pooling / transposed convolutions."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr",
                 num_groups=8, num_levels=5, is_segmentation=True, conv_padding=1, conv_upscale=2,
                 upsample="default", dropout_prob=0.1, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels,
                         final_sigmoid=final_sigmoid, basic_module=ResNetBlock, f_maps=f_maps,
                         layer_order=layer_order, num_groups=num_groups, num_levels=num_levels,
                         is_segmentation=is_segmentation, conv_padding=conv_padding,
                         conv_upscale=conv_upscale, upsample=upsample, dropout_prob=dropout_prob, is3d=False)


class UNet3D(AbstractUNet):
    """model.py:154-194."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr",
                 num_groups=8, num_levels=4, is_segmentation=True, conv_padding=1, conv_upscale=2,
                 upsample="default", dropout_prob=0.1, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels,
                         final_sigmoid=final_sigmoid, basic_module=DoubleConv, f_maps=f_maps,
                         layer_order=layer_order, num_groups=num_groups, num_levels=num_levels,
                         is_segmentation=is_segmentation, conv_padding=conv_padding,
                         conv_upscale=conv_upscale, upsample=upsample, dropout_prob=dropout_prob,
                         is3d=True)


class UNet2D(AbstractUNet):
    """model.py:283-320 — the same factory with 2D convolutions and pooling."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr",
                 num_groups=8, num_levels=4, is_segmentation=True, conv_padding=1, conv_upscale=2,
                 upsample="default", dropout_prob=0.1, **kwargs):
        super().__init__(in_channels=in_channels, out_channels=out_channels,
                         final_sigmoid=final_sigmoid, basic_module=DoubleConv, f_maps=f_maps,
                         layer_order=layer_order, num_groups=num_groups, num_levels=num_levels,
                         is_segmentation=is_segmentation, conv_padding=conv_padding,
                         conv_upscale=conv_upscale, upsample=upsample, dropout_prob=dropout_prob,
                         is3d=False)


def get_model(model_config):
    """model.py:362-364 — the YAML trainer's factory: ``{'name': 'UNet3D', ...}`` -> ``UNet3D(**config)``."""
    classes = {c.__name__: c for c in (UNet3D, UNet2D, ResidualUNet3D, ResidualUNetSE3D, ResidualUNet2D)}
    name = model_config["name"]
    if name not in classes:
        raise RuntimeError(f"Unsupported dataset class: {name}")     # utils.get_class's message (utils.py)
    return classes[name](**model_config)


if PreTrainedModel is not None:

    class UNet3DForMedicalSegmentationConfig(PretrainedConfig):
        """model/unet3d/UNet3D.py:18-59."""

        def __init__(self, in_channels=1, out_channels=1, final_sigmoid=True, basic_module="DoubleConv",
                     f_maps=64, layer_order="gcr", num_groups=8, num_levels=4, is_segmentation=True,
                     conv_kernel_size=3, pool_kernel_size=2, conv_padding=1, conv_upscale=2,
                     upsample="default", dropout_prob=0.1, is3d=True, loss_config: dict = None,
                     unet_type="UNet3D", **kwargs):
            super().__init__(**kwargs)
            self.in_channels, self.out_channels = in_channels, out_channels
            self.final_sigmoid, self.basic_module = final_sigmoid, basic_module
            self.f_maps, self.layer_order, self.num_groups = f_maps, layer_order, num_groups
            self.num_levels, self.is_segmentation = num_levels, is_segmentation
            self.conv_kernel_size, self.pool_kernel_size = conv_kernel_size, pool_kernel_size
            self.conv_padding, self.conv_upscale = conv_padding, conv_upscale
            self.upsample, self.dropout_prob, self.is3d = upsample, dropout_prob, is3d
            self.unet_type = unet_type
            self.loss_config = loss_config or {"loss": {"name": "BCEDiceLoss"}}

    @dataclass
    class UNet3DForMedicalSegmentationOutput(ModelOutput):
        """UNet3D.py:62-66."""
        loss: Optional[torch.FloatTensor] = None
        logits: Optional[torch.FloatTensor] = None
        labels: Optional[torch.LongTensor] = None

    class UNet3DForMedicalSegmentation(PreTrainedModel):
        """UNet3D.py:69-154, quirks included: the activation is applied BEFORE the loss
        criterion (so the default BCEDiceLoss sees a sigmoid twice) and ``logits`` carries the
        activated tensor."""
        config_class = UNet3DForMedicalSegmentationConfig

        def __init__(self, config):
            super().__init__(config)
            nets = {"UNet3D": UNet3D, "ResidualUNet3D": ResidualUNet3D, "ResidualUNetSE3D": ResidualUNetSE3D}
            if config.unet_type not in nets:                                 # UNet3D.py:72-118
                raise NotImplementedError(f"unet_type={config.unet_type!r} (the reference knows {sorted(nets)})")
            net_cls = nets[config.unet_type]
            self.model = net_cls(in_channels=config.in_channels, out_channels=config.out_channels,
                                final_sigmoid=config.final_sigmoid, f_maps=config.f_maps,
                                layer_order=config.layer_order, num_groups=config.num_groups,
                                num_levels=config.num_levels, is_segmentation=config.is_segmentation,
                                conv_padding=config.conv_padding, conv_upscale=config.conv_upscale,
                                upsample=config.upsample, dropout_prob=config.dropout_prob)
            if config.is_segmentation and config.final_sigmoid:
                self.activation = nn.Sigmoid()
            elif config.is_segmentation:
                self.activation = nn.Softmax(dim=1)
            else:
                self.activation = None
            self.loss_criterion = get_loss_criterion(config.loss_config)

        def _init_weights(self, module):
            return

        def forward(self, volume, target=None, weight=None):
            out = self.model(volume)
            activated = self.activation(out) if self.activation is not None else out
            loss = None
            if target is not None:
                loss = (self.loss_criterion(activated, target) if weight is None
                        else self.loss_criterion(activated, target, weight))
            return UNet3DForMedicalSegmentationOutput(loss=loss, logits=activated, labels=target)
