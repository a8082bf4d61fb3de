"""Drop-in 2D U-Net classes (same names, constructor arguments, attribute names and therefore the
same ``state_dict`` keys as model/unet2d/layers.py and model/unet2d/unet.py of the reference), with
the arithmetic executed by the sm_100a kernels.

The ``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``nn.BatchNorm2d`` children are parameter containers
only (PyTorch's default initialisation is thereby identical to the reference); ``UNet.forward`` is
one fused autograd node (engine.run_network).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import nn

from . import engine
from .losses import BCEWithLogitsLoss, CrossEntropyLoss, F1Loss, IoULoss, SegmentationLoss  # noqa: F401

try:  # HF-style wrappers keep the reference's base classes when transformers is importable
    from transformers import PretrainedConfig, PreTrainedModel
    from transformers.utils import ModelOutput
except Exception:  # pragma: no cover - transformers is part of the image
    PretrainedConfig = PreTrainedModel = ModelOutput = None


class DoubleConvolution(nn.Module):
    """layers.py:103-133 — two 3x3 convolutions (padding 1, bias), each followed by ReLU."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.first = nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1)
        self.act1 = nn.ReLU()
        self.second = nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1)
        self.act2 = nn.ReLU()

    def forward(self, x: torch.Tensor):
        return engine.run_block(self, engine.unet2d_specs(self), x)


class unetConv2(nn.Module):
    """layers.py:8-46 — n x (Conv -> [BatchNorm2d] -> ReLU), Kaiming-normal initialised."""

    def __init__(self, in_size, out_size, is_batchnorm, n=2, ks=3, stride=1, padding=1):
        super().__init__()
        if ks != 3 or stride != 1 or padding != 1:
            raise NotImplementedError("unetConv2 on the B200 path supports ks=3, stride=1, padding=1")
        self.n, self.ks, self.stride, self.padding = n, ks, stride, padding
        self.is_batchnorm = is_batchnorm
        for i in range(1, n + 1):
            layers = [nn.Conv2d(in_size, out_size, ks, stride, padding)]
            if is_batchnorm:
                layers.append(nn.BatchNorm2d(out_size))
            layers.append(nn.ReLU(inplace=True))
            setattr(self, "conv%d" % i, nn.Sequential(*layers))
            in_size = out_size
        for m in self.modules():  # init_weights(m, 'kaiming'), model/unet2d/init_weights.py:29-38
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.normal_(m.weight.data, 1.0, 0.02)
                nn.init.constant_(m.bias.data, 0.0)

    def _specs(self):
        out = []
        for i in range(1, self.n + 1):
            seq = getattr(self, "conv%d" % i)
            if self.is_batchnorm:
                out.append(engine.LayerSpec("cbr", seq[0], batchnorm=seq[1]))
            else:
                out.append(engine.LayerSpec("cr", seq[0]))
        return out

    def forward(self, inputs):
        return engine.run_block(self, self._specs(), inputs)


class DownSample(nn.Module):
    """layers.py:136-150 — 2x2 max-pool."""

    def __init__(self):
        super().__init__()
        self.pool = nn.MaxPool2d(2)

    def forward(self, x: torch.Tensor):
        return _PoolFunction.apply(x)


class _PoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        from . import ops
        xc = ops.from_nchw(x)
        ctx.xc = xc
        return ops.to_nchw(ops.maxpool_fwd(xc, 1), False)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        return ops.to_nchw(ops.maxpool_bwd(ctx.xc, ops.from_nchw(g), 1), False)


class UpSample(nn.Module):
    """layers.py:153-168 — ConvTranspose2d(kernel 2, stride 2)."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_channels, out_channels, kernel_size=2, stride=2)

    def forward(self, x: torch.Tensor):
        return _UpFunction.apply(x, self.up.weight, self.up.bias)


class _UpFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        from . import ops
        xc = ops.from_nchw(x)
        fwd, bwd = ops.convt_repack_weight(w)
        y = ops.convt_fwd(xc, fwd, None if b is None else b.detach().float().contiguous())
        ctx.xc, ctx.bwd, ctx.has_b = xc, bwd, b is not None
        return ops.to_nchw(y, False)

    @staticmethod
    def backward(ctx, g):
        from . import ops
        gc = ops.from_nchw(g)
        dx = ops.to_nchw(ops.convt_bwd_data(gc, ctx.bwd), False)
        return dx, ops.convt_bwd_weight(ctx.xc, gc), ops.bias_grad(gc) if ctx.has_b else None


class CropAndConcat(nn.Module):
    """layers.py:171-192 — centre-crop the skip to x and concatenate [x, skip]. Inside ``UNet`` this
    is fused away (both producers write into one buffer); stand-alone it is a plain copy."""

    def forward(self, x: torch.Tensor, contracting_x: torch.Tensor):
        h, w = x.shape[2], x.shape[3]
        sh, sw = contracting_x.shape[2], contracting_x.shape[3]
        top, left = int(round((sh - h) / 2.0)), int(round((sw - w) / 2.0))
        return torch.cat([x, contracting_x[:, :, top:top + h, left:left + w]], dim=1)


class UNet(nn.Module):
    """model/unet2d/unet.py:42-128 — 4 down (64..512), bottleneck 1024, 4 up, 1x1 head; logits out."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        widths = [64, 128, 256, 512]
        self.down_conv = nn.ModuleList(
            [DoubleConvolution(i, o) for i, o in zip([in_channels] + widths[:-1], widths)])
        self.down_sample = nn.ModuleList([DownSample() for _ in range(4)])
        self.middle_conv = DoubleConvolution(512, 1024)
        ups = [(1024, 512), (512, 256), (256, 128), (128, 64)]
        self.up_sample = nn.ModuleList([UpSample(i, o) for i, o in ups])
        self.up_conv = nn.ModuleList([DoubleConvolution(i, o) for i, o in ups])
        self.concat = nn.ModuleList([CropAndConcat() for _ in range(4)])
        self.final_conv = nn.Conv2d(64, out_channels, kernel_size=1)

    def forward(self, images: torch.Tensor):
        return engine.run_network(self, "unet2d", images)


if PreTrainedModel is not None:

    @dataclass
    class UNetModelOutput(ModelOutput):
        """unet.py:1156-1160."""
        loss: Optional[torch.FloatTensor] = None
        logits: Optional[torch.FloatTensor] = None
        labels: Optional[torch.LongTensor] = None

    class UNetConfig(PretrainedConfig):
        """unet.py:1163-1175."""

        def __init__(self, in_channels=1, out_channels=1, unet_type="UNet", **kwargs):
            super().__init__(**kwargs)
            self.in_channels = in_channels
            self.out_channels = out_channels
            self.unet_type = unet_type
            self.label_names = "labels"
            self.main_input_name = "images"
            self.keys_to_ignore_at_inference = ["labels"]

    class UNetModel(PreTrainedModel):
        """unet.py:1178-1213 — ``forward(images, labels) -> UNetModelOutput(loss, logits, labels)``.
        BCE-with-logits for one output channel, cross-entropy for several; both are single-pass
        reductions on the device."""
        config_class = UNetConfig

        def __init__(self, config: UNetConfig):
            super().__init__(config)
            if config.unet_type == "UNet":
                self.unet = UNet(config.in_channels, config.out_channels)
                self.criterion = CrossEntropyLoss() if config.out_channels > 1 else BCEWithLogitsLoss()
            elif config.unet_type in ("UNet_3Plus", "UNet_3Plus_DeepSup"):       # unet.py:1189-1196
                from . import unet3plus
                self.unet = getattr(unet3plus, config.unet_type)(config.in_channels, config.out_channels)
                self.criterion = SegmentationLoss()
            else:
                raise NotImplementedError(f"unet_type={config.unet_type!r} (the reference knows UNet, "
                                          "UNet_3Plus and UNet_3Plus_DeepSup, unet.py:1166-1168)")

        def _init_weights(self, module):  # keep PyTorch's default init, as the reference does
            return

        def forward(self, images: torch.Tensor, labels: Optional[torch.Tensor] = None):
            logits = self.unet(images)
            loss = None
            if isinstance(logits, tuple):          # deep supervision: criterion summed over d1..d5 (:1203-1208)
                if labels is not None:
                    loss = sum(self.criterion(d, labels) for d in logits)
                logits = logits[0]
            elif labels is not None:
                loss = self.criterion(logits, labels)
            return UNetModelOutput(loss=loss, logits=logits, labels=labels)
