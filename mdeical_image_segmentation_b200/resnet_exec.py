"""Whole-network executor of the residual U-Nets (``ResidualUNet3D`` / ``ResidualUNetSE3D``,
model/unet3d/model.py:197-280) — SURVEY.md §8(f) rank 2.

Topology (buildingblocks.py:255-325, 442-550, 604-639): encoder level = [2x max-pool] -> ResNetBlock; decoder
level = ConvTranspose(k3, s2, p1, no bias) + nearest resize to the skip's size -> **summation** joining ->
ResNetBlock(out, out).  ResNetBlock: ``residual = conv1(x)`` (1x1x1 conv with bias when the channel count
changes, else identity) -> ``conv2`` (SingleConv, e.g. GroupNorm-conv-ReLU) -> ``conv3`` (the same order without
its non-linearity) -> ``relu(out + residual)``.

Everything runs on kernels that already exist on the path — the 1x1x1 convolution is the 1-tap case of the
tcgen05 shift-GEMM, the learned up-sampling is the parity-plane transposed convolution of the ``deconv``
decoders — plus ``b200_add_relu`` for the residual join.  ReLU backward of a block output is applied where its
gradients meet: by the max-pool backward that also sums in the skip gradient (encoders), by ``b200_relu_bwd``
(decoders, bottom level), by the head backward (last block).
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import nn

from . import engine, ops
from .ops import CL


def _specs(block, order: str):
    n_order = order
    for ch in "rel":
        n_order = n_order.replace(ch, "")
    return engine.single_conv_spec(block.conv2, order), engine.single_conv_spec(block.conv3, n_order)


def _block_fwd(block, order: str, x: CL, training: bool, want: bool):
    """-> (relu(conv3(conv2(res)) + res), tape record)."""
    t1: Optional[list] = [] if want else None
    t2: Optional[list] = [] if want else None
    t3: Optional[list] = [] if want else None
    if isinstance(block.conv1, nn.Identity):
        res = x
    else:
        res, _, _ = engine.single_conv_fwd(x, engine.LayerSpec("c", block.conv1), training, t1)
        if res.zpad or not res.is_full():
            raise RuntimeError("residual branch must be a dense tensor")
    s2, s3 = _specs(block, order)
    out, _, _ = engine.single_conv_fwd(res, s2, training, t2)
    out, _, _ = engine.single_conv_fwd(out, s3, training, t3)
    y = ops.add_relu(out, res)
    se_rec = None
    if hasattr(block, "se_module"):
        y, se_rec = _se_fwd(block.se_module, y, want)
    return y, (t1, t2, t3, isinstance(block.conv1, nn.Identity), se_rec)


def _se_fwd(se, y: CL, want: bool):
    """ChannelSpatialSELayer3D (se.py:100-113) on a block output: z = y * max(cSE gate, sSE gate)."""
    cse, sse = se.cSE, se.sSE
    squeeze = ops.se_channel_means(y)                                    # (n, c), se.py:46
    with torch.enable_grad():                                             # two tiny FCs: library GEMMs + autograd
        sq = squeeze.detach().requires_grad_(want)
        hidden = torch.relu(torch.nn.functional.linear(sq, cse.fc1.weight, cse.fc1.bias))
        a = torch.sigmoid(torch.nn.functional.linear(hidden, cse.fc2.weight, cse.fc2.bias))
    a_val = a.detach().float().contiguous()
    bias = None if sse.conv.bias is None else sse.conv.bias.detach().float().contiguous()
    s = ops.head_fwd(y, sse.conv.weight, bias, True)                      # sSE logits (n, 1, d, h, w), se.py:88
    z = ops.se_gate_fwd(y, a_val, s)
    return z, ((y, a, sq, a_val, s, se) if want else None)


def _se_bwd(g: CL, rec, grads: engine.Grads) -> CL:
    """g = d(loss)/dz masked by z > 0 (== y > 0).  Returns d(loss)/d(out + residual), masked."""
    y, a, sq, a_val, s, se = rec
    cse, sse = se.cSE, se.sSE
    da, ds = ops.se_gate_bwd_reduce(y, a_val, s, g)
    params = [cse.fc1.weight, cse.fc1.bias, cse.fc2.weight, cse.fc2.bias]
    dsq, *dps = torch.autograd.grad(a, [sq] + params, grad_outputs=da.to(a.dtype), allow_unused=True)
    for p, dp in zip(params, dps):
        if dp is not None:
            engine._acc(grads, p, dp.float())
    dsse, dw, db = ops.head_bwd(y, sse.conv.weight, ds, mask_relu=True)
    engine._acc(grads, sse.conv.weight, dw)
    if sse.conv.bias is not None:
        engine._acc(grads, sse.conv.bias, db)
    dmean = (dsq.float() / float(y.spatial)).contiguous()
    return ops.se_gate_bwd_apply(y, a_val, s, g, dmean, dsse)


def _block_bwd(g: CL, rec, grads: engine.Grads, need_dx: bool = True) -> Optional[CL]:
    """g = d(loss)/d(out + residual), i.e. already masked by the block's final ReLU.  g is only read."""
    t1, t2, t3, identity, se_rec = rec
    if se_rec is not None:
        g = _se_bwd(g, se_rec, grads)
    d, _, _ = engine.single_conv_bwd(g, t3, grads)            # conv3 has no ReLU: g is not clobbered
    d, _, _ = engine.single_conv_bwd(d, t2, grads)
    d = ops.add_(d, g)                                        # + the skip connection's share
    if identity:
        return d
    dx, _, _ = engine.single_conv_bwd(d, t1, grads, need_dx=need_dx)
    return dx


def forward(net, x_in: torch.Tensor, training: bool, tape: Optional[list]) -> CL:
    is3d, order = net.is3d, net.layer_order
    want = tape is not None
    x = ops.from_nchw(x_in, pad=True)
    st = {"enc": [], "dec": [], "outs": [], "pool_in": []}
    skips: List[CL] = []
    for i, enc in enumerate(net.encoders):
        if enc.pooling is not None:
            st["pool_in"].append(x)
            x = ops.maxpool_fwd(x, 2 if is3d else 1)
        x, rec = _block_fwd(enc.basic_module, order, x, training, want)
        st["enc"].append(rec)
        skips.append(x)
    for j, dec in enumerate(net.decoders):
        skip = skips[len(skips) - 2 - j]
        ct = dec.upsampling.upsample.conv_transposed
        wp_fwd, wp_bwd, kd = ops.deconv_repack_weight(ct.weight)
        planes = ops.deconv_fwd(x, wp_fwd, kd, impl=engine.IMPL)
        up = ops.deconv_resize_fwd(planes, x.n, kd, (skip.d, skip.h, skip.w))
        joined = ops.add_(up, skip)                           # buildingblocks.py:549 (summation joining)
        y, rec = _block_fwd(dec.basic_module, order, joined, training, want)
        st["dec"].append((x, ct, wp_bwd, kd, rec))
        st["outs"].append(y)
        x = y
    if want:
        st["skips"] = skips
        tape.append(("resnet", st))
        tape.append(("res_end", None, True))                  # the feature map is a ReLU output
    return x


def backward(net, tape: list, dfeat: CL, grads: engine.Grads, premasked: bool) -> None:
    """dfeat: gradient w.r.t. the last block's output, already masked by the head backward (premasked)."""
    is3d = net.is3d
    tape.pop()
    _, st = tape.pop()
    skips = st["skips"]
    n_enc = len(skips)
    dskips: List[Optional[CL]] = [None] * n_enc
    g = dfeat
    if not premasked:
        g, _ = ops.relu_bwd(g, st["outs"][-1] if st["outs"] else skips[-1], False, out=g)
    for j in range(len(net.decoders) - 1, -1, -1):
        x_in, ct, wp_bwd, kd, rec = st["dec"][j]
        dj = _block_bwd(g, rec, grads)                        # gradient w.r.t. (up + skip)
        dskips[n_enc - 2 - j] = dj
        d_planes = ops.deconv_resize_bwd(dj, x_in.dims, kd)
        with engine._OnSide(x_in.buf, d_planes.buf):
            engine._acc(grads, ct.weight, ops.deconv_bwd_weight(x_in, d_planes, tuple(ct.weight.shape), kd,
                                                               impl=engine.IMPL))
        d = ops.deconv_bwd_data(d_planes, wp_bwd, x_in.dims, kd, impl=engine.IMPL)
        g, _ = ops.relu_bwd(d, x_in, False, out=d)            # x_in = the block output below (a ReLU output)
    # g is now the masked gradient of the deepest encoder's output
    for i in range(n_enc - 1, -1, -1):
        d = _block_bwd(g, st["enc"][i], grads, need_dx=i > 0)
        if i > 0:
            pre = st["pool_in"][i - 1]                        # output of encoder i-1: pooled AND used as a skip
            g = ops.maxpool_bwd(pre, d, 2 if is3d else 1, add=dskips[i - 1], mask_relu=True)
