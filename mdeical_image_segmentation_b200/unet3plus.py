"""UNet 3+ (model/unet2d/unet.py:136-446) on the B200 path — SURVEY.md §8(f) rank 1, the default model of
the reference's ``train.py`` (train.py:101-105).

Same constructor, attribute names and therefore ``state_dict`` keys as the reference.  The network is a
small DAG: five ``unetConv2`` encoder levels (64..1024 channels) and four decoder stages, each of which
gathers ALL five scales (max-pooled encoder maps, the same-scale encoder map, bilinearly up-sampled deeper
decoder maps), gives every one a 3x3 conv + BatchNorm + ReLU to 64 channels, concatenates them (320) and
fuses them with one more conv + BN + ReLU.  On this path

* every 3x3 convolution runs on the tcgen05 implicit-GEMM kernels (conv -> BN statistics in the epilogue ->
  one normalise + ReLU pass), the five branch results are written straight into their 64-channel slice of
  the 320-channel concat buffer (no ``torch.cat``),
* the k x k max-pools (k = 2, 4, 8) and the bilinear x2..x16 resizes are HBM-bound kernels
  (``b200_maxpoolk_*``, ``b200_bilinear2d_*``) whose backward kernels accumulate straight into the gradient
  of the tensor they read (an encoder map feeds up to five branches),
* the 3x3 output conv (320 -> n_classes) is padded to 16 output channels so it too runs on tensor cores.

``UNet_3Plus_DeepSup`` (unet.py:454-787) adds the deep-supervision heads: a 3x3 conv to n_classes on every
decoder scale (hd1..hd4 and the bottleneck hd5), bilinearly up-sampled to the input size; it returns the
tuple (d1, .., d5) of logits that ``UNetModel.forward`` (unet.py:1203-1208) sums the criterion over.  The
class-guided variant ``UNet_3Plus_DeepSup_CGM`` (unet.py:795-1153) adds the classification branch on the
bottleneck (dropout, 1x1 conv to 2 classes, global max, sigmoid, arg-max) whose 0/1 result gates the five
maps per sample before a sigmoid (``b200_cgm_gate``, ``b200_gate_sigmoid_*``); it returns probabilities.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F
from torch import nn

from . import engine, ops
from .ops import CL
from .unet2d import unetConv2

FILTERS = [64, 128, 256, 512, 1024]


def _kaiming(m: nn.Module) -> None:
    """init_weights(m, 'kaiming') — model/unet2d/init_weights.py:29-38."""
    if isinstance(m, nn.Conv2d):
        nn.init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0.0)


def _branch_name(level: int, stage: int) -> str:
    """Attribute prefix of the branch that brings scale ``level`` (1..5) to decoder ``stage`` (1..4)."""
    if level < stage:
        return f"h{level}_PT_hd{stage}"
    if level == stage:
        return f"h{level}_Cat_hd{stage}"
    return f"hd{level}_UT_hd{stage}"


class _UNet3PlusBody(nn.Module):
    """Encoder + four full-scale decoder stages shared by UNet_3Plus (unet.py:136-333) and
    UNet_3Plus_DeepSup (unet.py:454-662); modules are registered in the reference's order."""
    HEAD_SCALES = (1,)           # output head i reads hd{i+1} and is up-sampled by HEAD_SCALES[i]
    _b200_executor = True        # parallel.DataParallelUNet attaches its gradient reducer to such modules

    def __init__(self, in_channels=3, n_classes=1, feature_scale=4, is_deconv=True, is_batchnorm=True):
        super().__init__()
        if not is_batchnorm:
            raise NotImplementedError(f"{type(self).__name__}(is_batchnorm=False) is outside the B200 path")
        if n_classes > 16:
            raise NotImplementedError(f"{type(self).__name__} with more than 16 classes")
        self.is_deconv, self.in_channels = is_deconv, in_channels
        self.is_batchnorm, self.feature_scale = is_batchnorm, feature_scale
        f = FILTERS
        self.conv1 = unetConv2(in_channels, f[0], True)
        self.maxpool1 = nn.MaxPool2d(kernel_size=2)
        self.conv2 = unetConv2(f[0], f[1], True)
        self.maxpool2 = nn.MaxPool2d(kernel_size=2)
        self.conv3 = unetConv2(f[1], f[2], True)
        self.maxpool3 = nn.MaxPool2d(kernel_size=2)
        self.conv4 = unetConv2(f[2], f[3], True)
        self.maxpool4 = nn.MaxPool2d(kernel_size=2)
        self.conv5 = unetConv2(f[3], f[4], True)
        self.CatChannels, self.CatBlocks = f[0], 5
        self.UpChannels = self.CatChannels * self.CatBlocks
        for stage in (4, 3, 2, 1):          # same registration order as the reference -> same state_dict
            for level in range(1, 6):
                name = _branch_name(level, stage)
                if level < stage:
                    k = 2 ** (stage - level)
                    setattr(self, name, nn.MaxPool2d(k, k, ceil_mode=True))
                    cin = f[level - 1]
                elif level == stage:
                    cin = f[level - 1]
                else:
                    setattr(self, name, nn.Upsample(scale_factor=2 ** (level - stage), mode="bilinear"))
                    cin = f[4] if level == 5 else self.UpChannels
                setattr(self, name + "_conv", nn.Conv2d(cin, self.CatChannels, 3, padding=1))
                setattr(self, name + "_bn", nn.BatchNorm2d(self.CatChannels))
                setattr(self, name + "_relu", nn.ReLU(inplace=True))
            setattr(self, f"conv{stage}d_1", nn.Conv2d(self.UpChannels, self.UpChannels, 3, padding=1))
            setattr(self, f"bn{stage}d_1", nn.BatchNorm2d(self.UpChannels))
            setattr(self, f"relu{stage}d_1", nn.ReLU(inplace=True))
        self._make_heads(n_classes)
        for m in self.modules():
            _kaiming(m)

    def _make_heads(self, n_classes: int) -> None:
        raise NotImplementedError

    def _run(self, inputs: torch.Tensor):
        params = tuple(self.parameters())
        return _Net3PFunction.apply(self, inputs, *params)


class UNet_3Plus(_UNet3PlusBody):
    """unet.py:136-446.  ``forward(inputs) -> logits`` (the sigmoid is commented out in the reference)."""

    def _make_heads(self, n_classes):
        self.outconv1 = nn.Conv2d(self.UpChannels, n_classes, 3, padding=1)

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        return self._run(inputs)[0]


class UNet_3Plus_DeepSup(_UNet3PlusBody):
    """unet.py:454-787.  ``forward(inputs) -> (d1, d2, d3, d4, d5)``: logits of the five scales, each
    bilinearly up-sampled to the input size (:760-787; the sigmoids at :786 are commented out)."""
    HEAD_SCALES = (1, 2, 4, 8, 16)

    def _make_heads(self, n_classes):
        self.upscore6 = nn.Upsample(scale_factor=32, mode="bilinear")
        self.upscore5 = nn.Upsample(scale_factor=16, mode="bilinear")
        self.upscore4 = nn.Upsample(scale_factor=8, mode="bilinear")
        self.upscore3 = nn.Upsample(scale_factor=4, mode="bilinear")
        self.upscore2 = nn.Upsample(scale_factor=2, mode="bilinear")
        for i in range(1, 5):
            setattr(self, f"outconv{i}", nn.Conv2d(self.UpChannels, n_classes, 3, padding=1))
        self.outconv5 = nn.Conv2d(FILTERS[4], n_classes, 3, padding=1)

    def forward(self, inputs: torch.Tensor):
        return self._run(inputs)


class _GateSigmoid(torch.autograd.Function):
    """sigmoid(d * gate[sample]) — dotProduct (unet.py:1012-1017) + F.sigmoid (:1153); the gate is an arg-max
    result and carries no gradient."""

    @staticmethod
    def forward(ctx, d: torch.Tensor, gate: torch.Tensor):
        y = ops.gate_sigmoid_fwd(d, gate)
        ctx.save_for_backward(y, gate)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, gate = ctx.saved_tensors
        return ops.gate_sigmoid_bwd(y, dy, gate), None


class UNet_3Plus_DeepSup_CGM(_UNet3PlusBody):
    """unet.py:795-1153.  ``forward(inputs) -> (p1, .., p5)``: class-gated probabilities of the five scales.
    ``self.last_cls`` keeps the (n, 2) class probabilities of the last forward for inspection."""
    HEAD_SCALES = (1, 2, 4, 8, 16)

    def _make_heads(self, n_classes):
        UNet_3Plus_DeepSup._make_heads(self, n_classes)
        self.cls = nn.Sequential(nn.Dropout(p=0.5), nn.Conv2d(FILTERS[4], 2, 1), nn.AdaptiveMaxPool2d(1),
                                 nn.Sigmoid())
        self.last_cls = None

    def forward(self, inputs: torch.Tensor):
        *ds, gate = self._run(inputs)
        return tuple(_GateSigmoid.apply(d, gate) for d in ds)


# ===================================================================================== executor
def _spec(net: nn.Module, prefix: str) -> engine.LayerSpec:
    return engine.LayerSpec("cbr", getattr(net, prefix + "_conv"), batchnorm=getattr(net, prefix + "_bn"))


def _forward(net: "_UNet3PlusBody", images: torch.Tensor, training: bool, tape: Optional[dict]) -> Dict[int, CL]:
    if images.shape[2] % 16 or images.shape[3] % 16:
        raise ValueError("UNet_3Plus needs H and W divisible by 16 (the reference's torch.cat fails otherwise)")
    want = tape is not None
    x = ops.from_nchw(images, pad=True)
    h: Dict[int, CL] = {}          # scale level -> feature map (encoder h1..h5; h[5] is hd5)
    relu = False
    for lvl in range(1, 6):
        t = [] if want else None
        if lvl > 1:
            x = ops.maxpool_fwd(h[lvl - 1], 1)
        x, relu, _ = engine.block_fwd(x, getattr(net, f"conv{lvl}")._specs(), training, t, in_relu=relu)
        h[lvl] = x
        if want:
            tape[("enc", lvl)] = t
    hd: Dict[int, CL] = {5: h[5]}
    for stage in (4, 3, 2, 1):
        ref = h[stage]
        cat = CL.empty(ref.n, 1, ref.h, ref.w, net.UpChannels, ref.device)
        for level in range(1, 6):
            name = _branch_name(level, stage)
            if level < stage:
                src = ops.maxpoolk_fwd(h[level], 2 ** (stage - level))
            elif level == stage:
                src = h[level]
            else:
                src = ops.bilinear_fwd(hd[level], 2 ** (level - stage))
            t = [] if want else None
            engine.single_conv_fwd(src, _spec(net, name), training, t,
                                   out=cat.slice((level - 1) * net.CatChannels, net.CatChannels))
            if want:
                tape[("br", level, stage)] = t
        t = [] if want else None
        # the concat is five ReLU outputs: its gradient can carry their ReLU backward (in_relu)
        hd[stage], _, _ = engine.single_conv_fwd(
            cat, engine.LayerSpec("cbr", getattr(net, f"conv{stage}d_1"), batchnorm=getattr(net, f"bn{stage}d_1")),
            training, t, in_relu=True)
        if want:
            tape[("fuse", stage)] = t
    if want:
        tape["h"] = h
    return hd


def _pad16(w: torch.Tensor) -> torch.Tensor:
    """(n_classes, C, 3, 3) -> (16, C, 3, 3): the output conv runs on the tensor cores with N = 16."""
    return F.pad(w, (0, 0, 0, 0, 0, 0, 0, 16 - w.shape[0])) if w.shape[0] < 16 else w


class _Net3PFunction(torch.autograd.Function):
    """The whole network as one autograd node: returns one logits tensor per output head."""

    @staticmethod
    def forward(ctx, net: _UNet3PlusBody, images: torch.Tensor, *params):
        ops.require_cuda(images)
        need_grad = any(ctx.needs_input_grad)
        tape = {} if need_grad else None
        outs, heads = [], []
        with torch.no_grad(), engine.prepacked(net, need_grad):
            hd = _forward(net, images, net.training, tape)
            for i, scale in enumerate(net.HEAD_SCALES):
                oc = getattr(net, f"outconv{i + 1}")
                ncls = oc.weight.shape[0]
                wp_fwd, wp_bwd, k = ops.conv_repack_weight(_pad16(oc.weight.detach()), need_bwd=need_grad)
                bias = None if oc.bias is None else F.pad(oc.bias.detach().float(), (0, 16 - ncls)).contiguous()
                y = ops.conv_fwd(hd[i + 1], wp_fwd, k, bias, False, impl=engine.IMPL)
                if scale > 1:       # upscore{2..5}: the logits of a coarse scale, bilinearly to full size
                    y = ops.bilinear_fwd(y, scale)
                outs.append(ops.to_nchw(y.slice(0, ncls), False))
                heads.append((oc, wp_bwd, k, ncls, scale))
            gate = None
            if hasattr(net, "cls"):      # classification branch on hd5 (unet.py:1035-1038)
                drop, conv = net.cls[0], net.cls[1]
                keep, scale = None, 1.0
                if net.training and drop.p > 0:
                    keep = (torch.rand((hd[5].n, hd[5].h * hd[5].w, hd[5].c), device=images.device) >= drop.p)
                    keep, scale = keep.to(torch.uint8), 1.0 / (1.0 - drop.p)
                gate, net.last_cls = ops.cgm_gate(hd[5], conv.weight, conv.bias, keep, scale)
        ctx.net, ctx.tape, ctx.params = net, tape, params
        ctx.hd = hd if need_grad else None
        ctx.heads = heads
        if gate is None:
            return tuple(outs)
        ctx.mark_non_differentiable(gate)
        return tuple(outs) + (gate,)

    @staticmethod
    def backward(ctx, *dlogits):
        net, tape = ctx.net, ctx.tape
        if tape is None:
            raise RuntimeError("backward through a forward that ran without grad")
        grads: engine.Grads = {}
        # as engine._NetFunction.backward: parameter gradients go to the data-parallel reducer the moment they
        # are final, weight-gradient GEMMs run on the side stream next to the HBM-bound BatchNorm / resize /
        # pool backward kernels of the main chain
        engine._REDUCER = getattr(net, "_b200_reducer", None)
        engine._side_begin(dlogits[0].device if dlogits[0] is not None else ctx.hd[1].device)
        try:
            h, hd = tape["h"], ctx.hd
            # d(loss)/d hd1..hd5 from the output heads.  Every hd is a ReLU output, so its ReLU backward is
            # folded into the head's data gradient (mask = hd); a masked gradient may be summed with the raw
            # gradients that arrive later, because masking is idempotent.
            ghd: Dict[int, Optional[CL]] = {s: None for s in range(1, 6)}
            for i, (oc, wp_bwd, k, ncls, scale) in enumerate(ctx.heads):
                if dlogits[i] is None:
                    continue
                dy = ops.from_nchw(F.pad(dlogits[i].contiguous().float(), (0, 0, 0, 0, 0, 16 - ncls)))
                if scale > 1:
                    dy = ops.bilinear_bwd(dy, scale)
                dw, db = ops.conv_bwd_weight(hd[i + 1], dy, k, (16,) + tuple(oc.weight.shape[1:]), impl=engine.IMPL,
                                             bias_grad=True)
                engine._acc(grads, oc.weight, dw[:ncls].contiguous())
                if oc.bias is not None:
                    engine._acc(grads, oc.bias, db[:ncls].contiguous())
                ghd[i + 1] = ops.conv_bwd_data(dy, wp_bwd, k, mask=hd[i + 1], impl=engine.IMPL)
            dh: Dict[int, Optional[CL]] = {lvl: None for lvl in range(1, 6)}     # d(loss)/d h1..h5 (h5 = hd5)
            dh[5] = ghd[5]
            dhd: Dict[int, Optional[CL]] = {4: ghd[4], 3: ghd[3], 2: ghd[2]}
            cc = net.CatChannels
            for stage in (1, 2, 3, 4):
                g = ghd[1] if stage == 1 else dhd[stage]
                if g is None:
                    raise RuntimeError("UNet 3+ backward without a gradient for the full-resolution output")
                # only stage 1's gradient is fully masked (it has a single consumer, the head conv above);
                # deeper stages have accumulated raw gradients from the bilinear branches
                dcat, pm, _ = engine.single_conv_bwd(g, tape[("fuse", stage)], grads, premasked=(stage == 1))
                for level in range(1, 6):
                    d, _, _ = engine.single_conv_bwd(dcat.slice((level - 1) * cc, cc), tape[("br", level, stage)],
                                                     grads, premasked=pm)
                    if level < stage:
                        dh[level] = ops.maxpoolk_bwd(h[level], d, 2 ** (stage - level), add=dh[level])
                    elif level == stage:
                        dh[level] = d if dh[level] is None else ops.add_(dh[level], d)
                    elif level == 5:
                        dh[5] = ops.bilinear_bwd(d, 2 ** (level - stage), add=dh[5])
                    else:
                        dhd[level] = ops.bilinear_bwd(d, 2 ** (level - stage), add=dhd[level])
            d, pm = dh[5], False
            for lvl in range(5, 0, -1):
                d, _, _ = engine.block_bwd(d, tape[("enc", lvl)], 2, grads, need_dx=lvl > 1, premasked=pm)
                if lvl > 1:   # h[lvl-1] is a ReLU output: the pool backward masks the summed gradient by it
                    d = ops.maxpool_bwd(h[lvl - 1], d, 1, add=dh[lvl - 1], mask_relu=True)
                    pm = True
            engine._side_end(grads)
            if engine._REDUCER is not None:
                grads.update(engine._REDUCER.finish())
        finally:
            engine._REDUCER = None
            engine._side_end({})
        ctx.tape = ctx.hd = None
        return (None, None) + tuple(grads.get(p) if p.requires_grad else None for p in ctx.params)
