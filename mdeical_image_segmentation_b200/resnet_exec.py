"""Whole-network executor of the residual U-Nets (``ResidualUNet3D`` / ``ResidualUNetSE3D``,
model/unet3d/model.py:197-280) — SURVEY.md §8(f) rank 2.

Topology (buildingblocks.py:255-325, 442-550, 604-639): encoder level = [2x max-pool] -> ResNetBlock; decoder
level = ConvTranspose(k3, s2, p1, no bias) + nearest resize to the skip's size -> **summation** joining ->
ResNetBlock(out, out).  ResNetBlock: ``residual = conv1(x)`` (1x1x1 conv with bias when the channel count
changes, else identity) -> ``conv2`` (SingleConv, e.g. GroupNorm-conv-ReLU) -> ``conv3`` (the same order without
its non-linearity) -> ``relu(out + residual)``.

Everything runs on kernels that already exist on the path — the 1x1x1 convolution is the 1-tap case of the
tcgen05 shift-GEMM, the learned up-sampling is the parity-plane transposed convolution of the ``deconv``
decoders — plus ``b200_add_relu`` / ``b200_act_fwd`` for the residual join.  With a ReLU non-linearity the
backward of a block output is applied where its gradients meet: by the pool backward that also sums in the
skip gradient (encoders), by ``b200_relu_bwd`` (decoders, bottom level), by the head backward (last block).
LeakyReLU / ELU blocks ('l' / 'e' orders, ResNetBlock's own default ``cge``) cannot be folded that way (their
derivative is not a 0/1 mask): the same meeting points then run ``b200_act_bwd`` explicitly.

The bottom of the file runs the block classes stand-alone (``ResNetBlock.forward``, ``Decoder.forward``, the SE
layers' ``forward``; buildingblocks.py:314-325, 539-543, se.py) on fp32 NC[D]HW tensors through the same code.
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


def block_act(block):
    """(kind, param) of the block's final non-linearity (buildingblocks.py:306-311)."""
    nl = getattr(block, "non_linearity", None)
    if isinstance(nl, nn.LeakyReLU):
        return ops.ACT_LEAKY_RELU, float(nl.negative_slope)
    if isinstance(nl, nn.ELU):
        return ops.ACT_ELU, float(nl.alpha)
    return ops.ACT_RELU, 0.0


def net_is_relu(net) -> bool:
    return block_act(net.encoders[0].basic_module)[0] == ops.ACT_RELU


def _block_fwd(block, order: str, x: CL, training: bool, want: bool):
    """-> (act(conv3(conv2(res)) + res), tape record)."""
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
    kind, param = block_act(block)
    y = ops.add_relu(out, res) if kind == ops.ACT_RELU else ops.act_fwd(out, kind, param, add=res)
    se_rec = None
    if hasattr(block, "se_module"):
        y, se_rec = _se_fwd(block.se_module, y, want)
    return y, (t1, t2, t3, isinstance(block.conv1, nn.Identity), se_rec)


def _se_parts(se):
    """(cSE or None, sSE or None) of an scse / cse / sse module (buildingblocks.py:348-358)."""
    if hasattr(se, "cSE"):
        return se.cSE, se.sSE
    if hasattr(se, "fc1"):
        return se, None
    return None, se


def _se_fwd(se, y: CL, want: bool):
    """SE gate on a (non-negative) block output: z = y * max(cSE gate, sSE gate) (se.py:100-113); with only
    one of the two gates (``cse`` se.py:18-57 / ``sse`` se.py:60-97) the missing one is identically zero —
    both gates are sigmoids (> 0), so the maximum then IS the present gate."""
    cse, sse = _se_parts(se)
    a = sq = None
    if cse is not None:
        squeeze = ops.se_channel_means(y)                                    # (n, c), se.py:46
        with torch.enable_grad():                                             # two tiny FCs: library GEMMs + autograd
            sq = squeeze.detach().requires_grad_(want)
            hidden = torch.relu(torch.nn.functional.linear(sq, cse.fc1.weight, cse.fc1.bias))
            a = torch.sigmoid(torch.nn.functional.linear(hidden, cse.fc2.weight, cse.fc2.bias))
        a_val = a.detach().float().contiguous()
    else:
        a_val = torch.zeros((y.n, y.c), dtype=torch.float32, device=y.device)
    if sse is not None:
        bias = None if sse.conv.bias is None else sse.conv.bias.detach().float().contiguous()
        s = ops.head_fwd(y, sse.conv.weight, bias, True)                      # sSE logits (n, 1, d, h, w), se.py:88
    else:
        s = torch.full((y.n, 1, y.d, y.h, y.w), -1.0e4, dtype=torch.float32, device=y.device)   # sigmoid -> 0
    z = ops.se_gate_fwd(y, a_val, s)
    return z, ((y, a, sq, a_val, s, se) if want else None)


def _se_bwd(g: CL, rec, grads: engine.Grads) -> CL:
    """g = d(loss)/dz masked by z > 0 (== y > 0).  Returns d(loss)/d(out + residual), masked."""
    y, a, sq, a_val, s, se = rec
    cse, sse = _se_parts(se)
    da, ds = ops.se_gate_bwd_reduce(y, a_val, s, g)
    if cse is not None:
        params = [cse.fc1.weight, cse.fc1.bias, cse.fc2.weight, cse.fc2.bias]
        dsq, *dps = torch.autograd.grad(a, [sq] + params, grad_outputs=da.to(a.dtype), allow_unused=True)
        for p, dp in zip(params, dps):
            if dp is not None:
                engine._acc(grads, p, dp.float())
        dmean = (dsq.float() / float(y.spatial)).contiguous()
    else:
        dmean = torch.zeros((y.n, y.c), dtype=torch.float32, device=y.device)
    if sse is not None:
        dsse, dw, db = ops.head_bwd(y, sse.conv.weight, ds, mask_relu=True)
        engine._acc(grads, sse.conv.weight, dw)
        if sse.conv.bias is not None:
            engine._acc(grads, sse.conv.bias, db)
    else:
        dsse = CL(torch.zeros_like(y.buf))
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
            avg = getattr(enc, "pool_type", "max") == "avg"
            x = (ops.avgpool_fwd if avg else ops.maxpool_fwd)(x, 2 if is3d else 1)
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
        tape.append(("res_end", None, net_is_relu(net)))      # is the feature map a ReLU output?
    return x


def _out_act_bwd(g: CL, y: CL, act) -> CL:
    """Backward of a block's final non-linearity given its output y (in place on g)."""
    kind, param = act
    if kind == ops.ACT_RELU:
        return ops.relu_bwd(g, y, False, out=g)[0]
    return ops.act_bwd(g, y, kind, param, out=g)


def backward(net, tape: list, dfeat: CL, grads: engine.Grads, premasked: bool) -> None:
    """dfeat: gradient w.r.t. the last block's output, already masked by the head backward when the
    non-linearity is a ReLU (premasked)."""
    is3d = net.is3d
    tape.pop()
    _, st = tape.pop()
    skips = st["skips"]
    n_enc = len(skips)
    act = block_act(net.encoders[0].basic_module)
    relu = act[0] == ops.ACT_RELU
    dskips: List[Optional[CL]] = [None] * n_enc
    g = dfeat
    if not (premasked and relu):
        g = _out_act_bwd(g, st["outs"][-1] if st["outs"] else skips[-1], act)
    for j in range(len(net.decoders) - 1, -1, -1):
        x_in, ct, wp_bwd, kd, rec = st["dec"][j]
        dj = _block_bwd(g, rec, grads)                        # gradient w.r.t. (up + skip)
        dskips[n_enc - 2 - j] = dj
        d_planes = ops.deconv_resize_bwd(dj, x_in.dims, kd)
        with engine._OnSide(x_in.buf, d_planes.buf):
            engine._acc(grads, ct.weight, ops.deconv_bwd_weight(x_in, d_planes, tuple(ct.weight.shape), kd,
                                                               impl=engine.IMPL))
        d = ops.deconv_bwd_data(d_planes, wp_bwd, x_in.dims, kd, impl=engine.IMPL)
        g = _out_act_bwd(d, x_in, act)                        # x_in = the block output below
    # g is now the gradient of the deepest encoder's pre-activation sum
    for i in range(n_enc - 1, -1, -1):
        d = _block_bwd(g, st["enc"][i], grads, need_dx=i > 0)
        if i > 0:
            pre = st["pool_in"][i - 1]                        # output of encoder i-1: pooled AND used as a skip
            avg = getattr(net.encoders[i], "pool_type", "max") == "avg"
            g = (ops.avgpool_bwd if avg else ops.maxpool_bwd)(pre, d, 2 if is3d else 1, add=dskips[i - 1],
                                                              mask_relu=relu)
            if not relu:
                g = _out_act_bwd(g, pre, act)


# ===================================================================================== stand-alone blocks
def run_block_standalone(block, x: torch.Tensor) -> torch.Tensor:
    """``ResNetBlock(SE).forward`` (buildingblocks.py:314-325, 360-362) on an fp32 NC[D]HW tensor."""
    is3d = x.dim() == 5
    act = block_act(block)

    def run(inputs, want):
        xin = ops.from_nchw(inputs[0])
        y, rec = _block_fwd(block, block.order, xin, block.training, want)

        def bwd(dout):
            grads: engine.Grads = {}
            g = _out_act_bwd(ops.from_nchw(dout), y, act)      # y > 0 <=> (out + residual) > 0, SE or not
            dx = _block_bwd(g, rec, grads)
            return (ops.to_nchw(dx, is3d),), grads
        return ops.to_nchw(y, is3d), (bwd if want else None)
    return engine.run_standalone(run, (x,), tuple(block.parameters()))


def run_se_standalone(se, x: torch.Tensor) -> torch.Tensor:
    """``ChannelSELayer3D / SpatialSELayer3D / ChannelSpatialSELayer3D.forward`` (se.py:41-57, 75-97, 111-113).
    The fused gate needs x >= 0 (it is built for block outputs); a tensor with negative entries is refused."""
    if x.dim() != 5:
        raise ValueError("the 3D squeeze-and-excitation layers take (N, C, D, H, W) tensors")
    if bool((x < 0).any()):
        raise NotImplementedError("the fused SE gate (y * max(a, b)) of the B200 path needs a non-negative input "
                                  "(it runs on ReLU block outputs)")

    def run(inputs, want):
        y = ops.from_nchw(inputs[0])
        z, rec = _se_fwd(se, y, want)

        def bwd(dout):
            grads: engine.Grads = {}
            # the gate kernels fold the [y > 0] mask of a ReLU output (the layer is treated as following a ReLU,
            # as everywhere in the reference's networks): entries that are exactly 0 get gradient 0
            g = ops.from_nchw(dout)
            gm = ops.relu_bwd(g, y, False)[0]
            dy = _se_bwd(gm, rec, grads)
            return (ops.to_nchw(dy, True),), grads
        return ops.to_nchw(z, True), (bwd if want else None)
    return engine.run_standalone(run, (x,), tuple(se.parameters()))


def run_decoder_standalone(dec, encoder_features: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``Decoder.forward(encoder_features, x)`` (buildingblocks.py:539-550): up-sample x to the skip's size
    (nearest, or ConvTranspose k3 s2 p1 + nearest resize), concat ``(skip, x)`` or sum, basic module."""
    from .unet3d import DoubleConv
    is3d = x.dim() == 5
    ct = getattr(getattr(dec.upsampling, "upsample", None), "conv_transposed", None)
    is_double = isinstance(dec.basic_module, DoubleConv)
    order = dec.basic_module.order

    def run(inputs, want):
        skip_t, x_t = inputs
        skip = ops.from_nchw(skip_t)
        xin = ops.from_nchw(x_t)
        size = (skip.d, skip.h, skip.w)
        cx = ct.out_channels if ct is not None else xin.c
        if dec.concat:
            joined = CL.empty(skip.n, skip.d, skip.h, skip.w, skip.c + cx, skip.device)
            ops.copy_channels(skip, joined.slice(0, skip.c))
            up_dst = joined.slice(skip.c, cx)
        else:
            joined = up_dst = CL.empty(skip.n, skip.d, skip.h, skip.w, cx, skip.device)
        wp_bwd = kd = None
        if ct is not None:
            wp_fwd, wp_bwd, kd = ops.deconv_repack_weight(ct.weight)
            planes = ops.deconv_fwd(xin, wp_fwd, kd, impl=engine.IMPL)
            ops.deconv_resize_fwd(planes, xin.n, kd, size, out=up_dst)
        else:
            ops.upsample_nearest_fwd(xin, size, out=up_dst)
        if not dec.concat:
            if skip.c != cx:
                raise ValueError(f"summation joining of {skip.c} and {cx} channels")
            ops.add_(joined, skip)
        if is_double:
            tape: Optional[list] = [] if want else None
            y, y_relu, _ = engine.block_fwd(joined, engine.double_conv_specs(dec.basic_module, order),
                                            dec.basic_module.training, tape)
            rec = None
        else:
            tape = None
            y, rec = _block_fwd(dec.basic_module, order, joined, dec.basic_module.training, want)

        def bwd(dout):
            grads: engine.Grads = {}
            g = ops.from_nchw(dout)
            if is_double:
                dj, _, _ = engine.block_bwd(g, tape, 2, grads)
            else:
                dj = _block_bwd(_out_act_bwd(g, y, block_act(dec.basic_module)), rec, grads)
            d_skip = dj.slice(0, skip.c) if dec.concat else dj
            d_up = dj.slice(skip.c, cx) if dec.concat else dj
            if ct is not None:
                d_planes = ops.deconv_resize_bwd(d_up, xin.dims, kd)
                engine._acc(grads, ct.weight, ops.deconv_bwd_weight(xin, d_planes, tuple(ct.weight.shape), kd,
                                                                   impl=engine.IMPL))
                dx = ops.deconv_bwd_data(d_planes, wp_bwd, xin.dims, kd, impl=engine.IMPL)
            else:
                dx = ops.upsample_nearest_bwd(d_up, (xin.d, xin.h, xin.w))
            return (ops.to_nchw(d_skip, is3d), ops.to_nchw(dx, is3d)), grads
        return ops.to_nchw(y, is3d), (bwd if want else None)
    return engine.run_standalone(run, (encoder_features, x), tuple(dec.parameters()))
