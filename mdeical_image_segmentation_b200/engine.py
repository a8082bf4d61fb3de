"""Whole-network forward / backward executors for the 2D and 3D U-Nets.

The reference runs one stock ``nn.Module`` per op and lets autograd record ~100 nodes
(model/unet2d/unet.py:95-128, model/unet3d/model.py:125-151).  Here the module tree is only the
parameter container: each network is ONE ``torch.autograd.Function`` whose forward walks the tree,
launches the sm_100a kernels on channels-last bf16 buffers and keeps an explicit tape, and whose
backward replays the tape in reverse.  That gives the executor what autograd cannot express:

* skip-concat fusion — encoder outputs and up-sampled tensors are written by their producers
  directly into the two channel halves of one pre-allocated concat buffer (no ``torch.cat`` pass);
  in backward the skip half of the concat gradient is summed into the max-pool backward;
* conv epilogues carry bias + ReLU; norm kernels carry the ReLU that follows them;
* gradients of the fp32 master parameters come back in the reference's own layouts.
"""
from __future__ import annotations

import os

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops
from .ops import CL

Grads = Dict[nn.Parameter, torch.Tensor]
IMPL = 0  # B200_IMPL_AUTO; tests flip it to cross-check the tcgen05 and direct kernels
FUSE_SUM_MIN_K = 2304  # fuse a bias-gradient column sum into a dgrad epilogue only if taps * C_out >= this
FUSE_BIAS_IN_DGRAD = os.environ.get("B200_FUSE_BIAS_IN_DGRAD", "0") != "0"   # A/B knob (superseded by wgrad fusion)
FUSE_DOT = os.environ.get("B200_FUSE_DOT", "1") != "0"   # A/B knob: norm-backward sums in the dgrad epilogue


_REDUCER = None  # parallel.GradReducer of the network whose backward is running (data parallel)

# Weight gradients are off the critical path of backward (only the optimizer / all-reduce consume them),
# so inside a whole-network backward they are enqueued on a side stream: the HBM-bound kernels of the main
# chain (norm backward, pool / up-sampling backward) then run next to a tensor-core-bound weight-gradient
# kernel instead of after it, and each stream fills the other's tail waves.
WGRAD_STREAM = os.environ.get("B200_WGRAD_STREAM", "1") != "0"
_SIDE = {}          # device index -> side stream
_side_active = None  # side stream of the backward that is running, or None


class _OnSide:
    """Runs the body on the side stream, ordered after everything enqueued so far on the current stream;
    ``keep`` are tensors the body reads: the caching allocator must not recycle them early."""

    def __init__(self, *keep):
        self.keep = keep
        self.ctx = None

    def __enter__(self):
        side = _side_active
        if side is None:
            return self
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        side.wait_event(ev)
        for t in self.keep:
            if t is not None:
                t.record_stream(side)
        self.ctx = torch.cuda.stream(side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _side_begin(device: torch.device):
    global _side_active
    if not WGRAD_STREAM or device.type != "cuda":
        _side_active = None
        return
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx not in _SIDE:
        _SIDE[idx] = torch.cuda.Stream(device=device)
    _side_active = _SIDE[idx]


def _side_end(grads: "Grads"):
    """Order the current stream after the side stream; gradients allocated there are now used here."""
    global _side_active
    side, _side_active = _side_active, None
    if side is None:
        return
    cur = torch.cuda.current_stream()
    cur.wait_stream(side)
    for g in grads.values():
        if g is not None:
            g.record_stream(cur)


def _acc(grads: Grads, p: Optional[nn.Parameter], g: torch.Tensor) -> None:
    """Record the (final) gradient of a parameter; under data parallelism hand it to the bucketed
    all-reduce immediately so the transfer overlaps the rest of the backward pass."""
    if p is None or not p.requires_grad:
        return
    g = g.reshape(p.shape)
    if p in grads:
        raise RuntimeError("a parameter received two gradients: weight sharing is not supported")
    grads[p] = _REDUCER.on_grad(p, g) if _REDUCER is not None else g


# ===================================================================================== filter packs
# The bf16 [tap][n][k] packs of all conv filters of a network are produced by ONE launch at the start of its
# forward (b200_conv_repack_weights): the first forward records which (conv, padded c_in) pairs the executor
# asks for, later forwards pre-pack that plan.  A conv that is not in the plan falls back to its own launch.
PREPACK = os.environ.get("B200_PREPACK", "1") != "0"   # A/B knob
_PACKS = None          # id(conv) -> (c_in_pad, wp_fwd, wp_bwd, k) while a network forward is running
_PLAN_RECORD = None    # list being recorded during a first forward


class prepacked:
    """Context manager around a whole-network forward."""

    def __init__(self, net, need_bwd: bool):
        self.net, self.need_bwd = net, need_bwd

    def __enter__(self):
        global _PACKS, _PLAN_RECORD
        plan = getattr(self.net, "_b200_pack_plan", None) if PREPACK else None
        if plan:
            res = ops.conv_repack_weights([c.weight for c, _ in plan], [p for _, p in plan], self.need_bwd)
            _PACKS = {id(c): (p,) + r for (c, p), r in zip(plan, res)}
            _PLAN_RECORD = None
        else:
            _PACKS, _PLAN_RECORD = None, ([] if PREPACK else None)
        return self

    def __exit__(self, *exc):
        global _PACKS, _PLAN_RECORD
        if _PLAN_RECORD and exc[0] is None:
            self.net._b200_pack_plan = _PLAN_RECORD
        _PACKS = _PLAN_RECORD = None
        return False


# ===================================================================================== SingleConv
SUPPORTED_LETTERS = set("cgbrledD")
# activation letters of create_conv (buildingblocks.py:55-59): nn.LeakyReLU() default slope, nn.ELU() default alpha
ACT_OF_LETTER = {"r": (ops.ACT_RELU, 0.0), "l": (ops.ACT_LEAKY_RELU, 0.01), "e": (ops.ACT_ELU, 1.0)}


class LayerSpec:
    """One conv layer in the reference's ``layer_order`` mini-language (buildingblocks.py:14-113):
    ``conv`` plus optional ``groupnorm`` / ``batchnorm`` modules, applied in ``order``."""

    def __init__(self, order: str, conv: nn.Module, groupnorm: Optional[nn.GroupNorm] = None,
                 batchnorm: Optional[nn.Module] = None, dropout_p: float = 0.0):
        bad = set(order) - SUPPORTED_LETTERS
        if bad:
            raise ValueError(f"Unsupported layer type(s) {sorted(bad)}. MUST be one of "
                             "['b', 'g', 'r', 'l', 'e', 'c', 'd', 'D']")
        self.order, self.conv, self.groupnorm, self.batchnorm = order, conv, groupnorm, batchnorm
        self.dropout_p = float(dropout_p)


def _bn_moments(bn: nn.Module, x: CL, training: bool, in_sums: Optional[torch.Tensor] = None):
    """mean / rstd for BatchNorm; in training also updates the running statistics the way
    nn.BatchNorm does (momentum 0.1, unbiased running variance).  ``in_sums`` (fp64 [n][2][c],
    per-sample sum x / sum x^2 from the producing conv's epilogue) replaces the statistics pass."""
    if training or bn.running_mean is None:
        sums = in_sums.sum(0) if in_sums is not None else ops.bn_stats(x)
        m = float(x.npos)
        mean = sums[0] / m
        var = (sums[1] / m - mean * mean).clamp_(min=0.0)
        if training and bn.running_mean is not None:
            with torch.no_grad():
                mom = bn.momentum if bn.momentum is not None else 0.1
                unbiased = var * (m / max(m - 1.0, 1.0))
                bn.running_mean.mul_(1 - mom).add_(mean.to(bn.running_mean.dtype), alpha=mom)
                bn.running_var.mul_(1 - mom).add_(unbiased.to(bn.running_var.dtype), alpha=mom)
                bn.num_batches_tracked += 1
        rstd = torch.rsqrt(var + bn.eps)
        return mean.float().contiguous(), rstd.float().contiguous()
    mean = bn.running_mean.float().contiguous()
    rstd = torch.rsqrt(bn.running_var.float() + bn.eps).contiguous()
    return mean, rstd


def _first_norm_grads(norm_rec, conv: nn.Module, dw: torch.Tensor, dw_ones: torch.Tensor, grads: "Grads") -> None:
    """dgamma / dbeta of the norm in front of the first conv from that conv's weight gradient (see
    _plant_ones): a few hundred multiply-adds on (c_out, c_in, taps) tensors."""
    kind = norm_rec[0]
    norm = norm_rec[2]
    w = conv.weight.detach().float()
    dbeta = (w * dw_ones.unsqueeze(1)).flatten(2).sum(dim=(0, 2))                    # (c_in,)
    gamma = norm.weight.detach().float()
    beta = norm.bias.detach().float()
    num = (w * dw).flatten(2).sum(dim=(0, 2)) - beta * dbeta
    safe = gamma.abs() > 1e-12
    dgamma = torch.where(safe, num / torch.where(safe, gamma, torch.ones_like(gamma)), torch.zeros_like(num))
    if kind == "bn" and not norm_rec[7]:      # eval-mode BatchNorm: same affine map, same formulas
        pass
    _acc(grads, norm.weight, dgamma)
    _acc(grads, norm.bias, dbeta)


def _im2col_kp(x_real: CL, conv: nn.Module, c_col: Optional[int] = None) -> int:
    """K' = taps * c_in rounded up to 16 when the first layer's WEIGHT GRADIENT should run as one short-K
    GEMM over an im2col of the narrow input, else 0.  Measured (tools/first_layer_probe.py,
    profiles/r01_first_layer_probe.log): 3D 1->32 @128^3 0.72 -> 0.20 + 0.18 ms; the forward / data
    gradient lose (a 128-row tile with two MMAs is bound by the ~1700-cycle accumulator hand-shake, which
    the halo kernel amortises over 512-position super-tiles) and so does 2D, where the im2col write costs
    what the weight gradient saves."""
    w = conv.weight
    c_col = x_real.c if c_col is None else c_col      # real channels (+ the planted ones channel)
    if w.dim() != 5 or tuple(w.shape[2:]) != (3, 3, 3) or x_real.coff != 0 or c_col - w.shape[1] not in (0, 1):
        return 0
    kp = (27 * c_col + 15) // 16 * 16
    return kp if kp <= 64 else 0


def _im2col_weight_grad(dw2: torch.Tensor, w: torch.Tensor, c_col: Optional[int] = None) -> torch.Tensor:
    """(c_out, kp) with k' = tap * c_col + ci  ->  (c_out, c_col, kd, kh, kw) (the conv's own layout when c_col
    is its c_in)."""
    co, ci = w.shape[0], (w.shape[1] if c_col is None else c_col)
    taps = w[0, 0].numel()
    shape = (co, ci) + tuple(w.shape[2:])
    return dw2.reshape(co, -1)[:, :taps * ci].reshape(co, taps, ci).permute(0, 2, 1).reshape(shape).contiguous()


FIRST_NORM_FROM_WGRAD = os.environ.get("B200_FIRST_NORM_FROM_WGRAD", "1") != "0"   # A/B knob
FIRST_LAYER_DIRECT = os.environ.get("B200_FIRST_LAYER_DIRECT", "1") != "0"         # A/B knob


def _first_layer_direct(conv: nn.Module, c_real: int, x: CL) -> bool:
    """nn.Conv3d(1, 32, 3, padding=1) on a volume stored in channel 0 of a padded buffer: the CUDA-core forward of
    csrc/first_layer.cu (0.37 -> 0.30 ms at 128^3 x 2) instead of the zero-padded tensor-core one.  The weight
    gradient stays on the tensor cores (im2col + short-K GEMM)."""
    w = conv.weight
    return (FIRST_LAYER_DIRECT and c_real == 1 and x.coff == 0 and w.dim() == 5 and tuple(w.shape) == (32, 1, 3, 3, 3)
            and min(x.d, x.h, x.w) >= 2)


def _plant_ones(dst: CL, tape, i: int, letters: str, nxt: int, fuse_relu: bool, conv) -> int:
    """A layer that starts ``norm -> conv`` on the narrow network input (stored zero-padded to 16 channels)
    needs d(loss)/d(norm output) only for the norm's two parameter gradients.  With u = gamma * xhat + beta the
    conv's input,
        dbeta_c  = sum_pos du_c           = <w[:, c], dW1>            dW1[co, tap] = sum_q dy[q, co] [q + tap inside]
        dgamma_c = sum_pos du_c * xhat_c  = (<w[:, c], dW[:, c]> - beta_c * dbeta_c) / gamma_c
    and dW1 is exactly the weight gradient of an input channel that is 1 everywhere: it is planted in the first
    padding channel (its filter taps are zero, so the forward is untouched) and rides in the layer's own
    weight-gradient GEMM.  The data-gradient launch and the norm-backward pass of the first layer disappear
    (cfg3: 0.70 + 0.08 ms of a 27.9 ms step).  Returns the channel index, or -1 when not applicable."""
    if (not FIRST_NORM_FROM_WGRAD or tape is None or i != 0 or fuse_relu or nxt >= len(letters)
            or letters[nxt] != "c" or dst.zpad < 1 or dst.coff != 0):
        return -1
    if not _im2col_kp(CL(dst.buf, dst.c + 1, 0, dst.zpad - 1), conv, dst.c + 1):
        dst.buf[..., dst.c].fill_(1.0)      # (the im2col weight-gradient path synthesises the column instead)
    return dst.c


class UpFold:
    """A decoder concat (skip | nearest 1x/2x up-sampling of ``low``) whose up-sampled half [cs, c) has not been
    written: the GroupNorm that reads it normalises from ``low`` directly (ops.gn_apply_up) and its backward reads
    ``low`` again.  Any other reader calls materialise() first."""

    def __init__(self, cat: CL, cs: int, low: CL):
        self.cat, self.cs, self.low, self.done = cat, cs, low, False
        self.dot_is_y = False      # backward: the data-gradient statistics were correlated with the norm's output

    def materialise(self) -> None:
        if not self.done:
            ops.upsample_nearest_fwd(self.low, (self.cat.d, self.cat.h, self.cat.w),
                                     out=self.cat.slice(self.cs, self.cat.c - self.cs))
            self.done = True


def single_conv_fwd(x: CL, spec: LayerSpec, training: bool, tape: Optional[list],
                    out: Optional[CL] = None, in_relu: bool = False,
                    in_sums: Optional[torch.Tensor] = None, want_sums: bool = False,
                    up_src: Optional[UpFold] = None):
    """Runs one layer; ``out`` (optional) is the destination slice of its final result.
    ``in_relu`` says the input is a ReLU output (>= 0, zero exactly where its producer was clipped):
    the first op's backward then folds that producer's ReLU backward into its own output.
    ``in_sums`` / ``want_sums``: per-sample per-channel (sum, sum of squares) of the input / of the
    result (fp64 [n][2][c]), produced in conv epilogues so that a norm never re-reads its input for
    statistics.  Returns (y, y_is_relu_output, y_sums or None)."""
    letters = spec.order
    cur = x
    cur_sums = in_sums if (in_sums is not None and in_sums.shape[2] == x.c) else None
    if up_src is not None and not (letters[0] == "g" and cur_sums is not None):
        up_src.materialise()          # only a GroupNorm with known statistics can skip the up-sampled half
        up_src = None
    i = 0
    out_relu = False
    ones_ch = -1          # index of a constant-one channel planted in the padding of the conv's input
    while i < len(letters):
        ch = letters[i]
        fuse_relu = ch in "cgb" and i + 1 < len(letters) and letters[i + 1] == "r"
        nxt = i + (2 if fuse_relu else 1)
        dst = out if nxt >= len(letters) else None
        if ch == "c":
            conv = spec.conv
            w = conv.weight.detach()
            xin = cur
            if cur.zpad and _first_layer_direct(conv, cur.c, cur):
                bias = None if conv.bias is None else conv.bias.detach().float().contiguous()
                need = (letters[nxt] in "gb") if nxt < len(letters) else want_sums
                y, cur_sums = ops.conv3d_c1_fwd(cur, w, bias, fuse_relu, out=dst, stats=need)
                if tape is not None:
                    tape.append(("conv", cur.widened(), conv, None, (3, 3, 3), y if fuse_relu else None, cur.zpad,
                                 in_relu and i == 0, ones_ch))
                ones_ch = -1
                cur = y
                out_relu = fuse_relu
                i = nxt
                continue
            pk = _PACKS.get(id(conv)) if _PACKS is not None else None
            if pk is not None and pk[0] == w.shape[1] + cur.zpad and (pk[2] is not None or tape is None):
                wp_fwd, wp_bwd, k = pk[1], pk[2], pk[3]      # pre-packed (and channel-padded) by the one launch
                if cur.zpad:
                    xin = cur.widened()
            else:
                if cur.zpad:  # narrow network input stored zero-padded: pad the filter's c_in to match
                    xin = cur.widened()
                    if _PLAN_RECORD is not None:
                        _PLAN_RECORD.append((conv, int(w.shape[1] + cur.zpad)))
                    w = torch.nn.functional.pad(w, (0, 0) * (w.dim() - 2) + (0, cur.zpad))
                elif _PLAN_RECORD is not None:
                    _PLAN_RECORD.append((conv, int(w.shape[1])))
                wp_fwd, wp_bwd, k = ops.conv_repack_weight(w, need_bwd=tape is not None)
            bias = None if conv.bias is None else conv.bias.detach().float().contiguous()
            need = (letters[nxt] in "gb") if nxt < len(letters) else want_sums
            if need:
                y, cur_sums = ops.conv_fwd(xin, wp_fwd, k, bias, fuse_relu, out=dst, impl=IMPL, stats=True)
            else:
                y, cur_sums = ops.conv_fwd(xin, wp_fwd, k, bias, fuse_relu, out=dst, impl=IMPL), None
            if tape is not None:
                tape.append(("conv", xin, conv, wp_bwd, k, y if fuse_relu else None, cur.zpad,
                             in_relu and i == 0, ones_ch))
            ones_ch = -1
            cur = y
        elif ch == "g":
            gn = spec.groupnorm
            if cur_sums is not None:
                stats = ops.gn_stats_from_sums(cur_sums, cur.spatial, gn.num_groups, gn.eps)
            else:
                stats = ops.gn_stats(cur, gn.num_groups, gn.eps)
            cur_sums = None
            gamma = gn.weight.detach().float().contiguous()
            beta = gn.bias.detach().float().contiguous()
            if dst is None and cur.zpad:
                dst = cur.new_like_padded()
                ones_ch = _plant_ones(dst, tape, i, letters, nxt, fuse_relu, spec.conv)
            fold = up_src if i == 0 else None
            if fold is not None:
                y = ops.gn_apply_up(cur, fold.cs, fold.low, gn.num_groups, stats, gamma, beta, fuse_relu, out=dst)
            else:
                y = ops.gn_apply(cur, gn.num_groups, stats, gamma, beta, fuse_relu, out=dst)
            if tape is not None:
                tape.append(("gn", cur, gn, stats, gamma, y if fuse_relu else None, in_relu and i == 0, fold))
            cur = y
        elif ch == "b":
            bn = spec.batchnorm
            mean, rstd = _bn_moments(bn, cur, training, cur_sums)
            cur_sums = None
            gamma = bn.weight.detach().float().contiguous()
            beta = bn.bias.detach().float().contiguous()
            if dst is None and cur.zpad:
                dst = cur.new_like_padded()
                ones_ch = _plant_ones(dst, tape, i, letters, nxt, fuse_relu, spec.conv)
            y = ops.bn_apply(cur, mean, rstd, gamma, beta, fuse_relu, out=dst)
            if tape is not None:
                tape.append(("bn", cur, bn, mean, rstd, gamma, y if fuse_relu else None, training,
                             in_relu and i == 0))
            cur = y
        elif ch in "rle":
            # LeakyReLU / ELU (and a ReLU that follows neither a conv nor a norm): one element-wise pass; the
            # backward recovers the derivative from the stored output
            kind, param = ACT_OF_LETTER[ch]
            y = ops.act_fwd(cur, kind, param, out=dst)
            if tape is not None:
                tape.append(("act", kind, param, y))
            cur, cur_sums = y, None
            fuse_relu = ch == "r"
        elif ch in "dD":
            # nn.Dropout (element-wise) / nn.Dropout2d (whole channels per sample), buildingblocks.py:104-107:
            # identity in eval mode; the keep mask comes from torch's generator, the multiply is ours
            p = spec.dropout_p
            if training and p > 0.0:
                scale = 0.0 if p >= 1.0 else 1.0 / (1.0 - p)
                keep = chan = None
                if ch == "d":
                    keep = (torch.rand((cur.npos, cur.c), device=cur.device) >= p).to(torch.uint8)
                else:
                    chan = (torch.rand((cur.n, cur.c), device=cur.device) >= p).float()
                y = ops.scale_mask(cur, keep, chan, scale, out=dst)
                if tape is not None:
                    tape.append(("drop", keep, chan, scale))
                cur, cur_sums = y, None
            fuse_relu = False if (training and p > 0.0) else out_relu
        out_relu = fuse_relu
        i = nxt
    if out is not None and cur is not out:
        ops.copy_channels(cur, out)
        cur = out
    return cur, out_relu, cur_sums


def _consumer_plan(consumer, fuse_mask: bool):
    """What the kernel producing a gradient should reduce for the op that consumes it (``consumer`` =
    that op's tape record, or ("sum",) for a plain per-channel sum).  Returns (want_stats, dot) where
    dot is the tensor to correlate the gradient with (None: sum of squares, unused)."""
    if consumer is None:
        return False, None
    kind = consumer[0]
    if kind == "sum":
        return True, None
    relu_out = consumer[5] if kind in ("conv", "gn") else consumer[6] if kind == "bn" else None
    if relu_out is not None and not fuse_mask:
        return False, None       # the consumer's ReLU backward is not folded into this gradient
    if kind == "conv":           # its bias gradient rides in its own weight-gradient kernel (fused there)
        conv = consumer[2]
        return bool(FUSE_BIAS_IN_DGRAD and conv.bias is not None and conv.bias.requires_grad), None
    if kind in ("gn", "bn"):     # (sum g, sum g * x) with x = the norm's input
        if not FUSE_DOT:
            return False, None
        if kind == "bn" and not consumer[7]:
            return False, None   # frozen statistics: no batch terms, the reduction path handles dgamma
        return True, consumer[1]
    return False, None


FOLD_UPSAMPLE_BWD = os.environ.get("B200_FOLD_UPSAMPLE_BWD", "1") != "0"   # A/B knob
FOLD_UPSAMPLE_FWD = os.environ.get("B200_FOLD_UPSAMPLE_FWD", "1") != "0"   # A/B knob


class FoldedCatGrad:
    """Gradient of a decoder concat buffer whose up-sampled half was reduced to its source resolution inside
    the GroupNorm backward (b200_gn_bwd_fused_up): ``skip`` (full resolution) and ``low``."""
    __slots__ = ("skip", "low")

    def __init__(self, skip: CL, low: CL):
        self.skip, self.low = skip, low


def single_conv_bwd(dy: CL, recs: list, grads: Grads, need_dx: bool = True, premasked: bool = False,
                    dy_sums: Optional[torch.Tensor] = None, consumer=None, up_fold=None):
    """Backward of one layer: consumes the layer's tape records (in reverse); dy may be clobbered.
    ``premasked``: dy already carries the ReLU backward of this layer's final ReLU (its producer fused
    it).  ``dy_sums``: fp64 [n][2][c] reductions of dy that its producer computed for this layer's last
    op (see _consumer_plan).  ``consumer``: tape record of the op that will receive dx.
    ``up_fold`` = (cs, low_dims): the layer's input is a decoder concat whose channels [cs, c) are an exact
    nearest 1x/2x up-sampling of a (low_dims) tensor — when the layer starts with a GroupNorm whose sums came
    from the data-gradient epilogue, dx is returned as a FoldedCatGrad (up-sampling backward folded in).
    Returns (dx, dx_premasked, dx_sums)."""
    out_premasked = False
    sums = dy_sums
    for idx in range(len(recs) - 1, -1, -1):
        rec = recs[idx]
        kind = rec[0]
        last = idx == 0
        pm = premasked and idx == len(recs) - 1
        nxt_consumer = recs[idx - 1] if idx > 0 else consumer
        if kind == "conv":
            _, x_in, conv, wp_bwd, k, y_relu, zpad, mask_in, ones_ch = rec
            # first layer "norm -> conv" on the narrow input: the norm's parameter gradients come out of this
            # layer's weight gradient (see _plant_ones) when nobody needs the data gradient itself
            from_wgrad = (ones_ch >= 0 and idx == 1 and not need_dx and conv.weight.requires_grad
                          and recs[0][0] in ("gn", "bn"))
            has_bias = conv.bias is not None and conv.bias.requires_grad
            db, db_in_wgrad = None, False
            if y_relu is not None and not pm:
                dy, db = ops.relu_bwd(dy, y_relu, has_bias, out=dy)
            elif has_bias and sums is not None:
                db = sums[:, 0, :].sum(0).float()
            elif has_bias and conv.weight.requires_grad:
                db_in_wgrad = True      # column sums of dy come out of the weight-gradient kernel
            elif has_bias:
                db = ops.bias_grad(dy)
            sums = None
            c_real = x_in.c - zpad
            c_col = c_real + (1 if from_wgrad else 0)     # channels the im2col has to carry
            kp = _im2col_kp(CL(x_in.buf, c_col, x_in.coff, x_in.c - c_col), conv, c_col) if zpad else 0
            dw_ones = None
            with _OnSide(x_in.buf, dy.buf, db):
                if conv.weight.requires_grad and kp:
                    a = ops.im2col(CL(x_in.buf, c_real, x_in.coff, x_in.c - c_real), 3, kp, ones=from_wgrad)
                    shape2 = (conv.weight.shape[0], kp, 1, 1, 1)
                    if db_in_wgrad:
                        dw2, db = ops.conv_bwd_weight(a, dy, (1, 1, 1), shape2, impl=IMPL, bias_grad=True)
                    else:
                        dw2 = ops.conv_bwd_weight(a, dy, (1, 1, 1), shape2, impl=IMPL)
                    dw_all = _im2col_weight_grad(dw2, conv.weight, c_col)          # (c_out, c_col, kd, kh, kw)
                    dw = dw_all[:, :c_real].contiguous() if from_wgrad else dw_all
                    dw_ones = dw_all[:, c_real] if from_wgrad else None
                    _acc(grads, conv.weight, dw)
                elif conv.weight.requires_grad:
                    wshape = list(conv.weight.shape)
                    wshape[1] += zpad
                    if db_in_wgrad:
                        dw, db = ops.conv_bwd_weight(x_in, dy, k, tuple(wshape), impl=IMPL, bias_grad=True)
                    else:
                        dw = ops.conv_bwd_weight(x_in, dy, k, tuple(wshape), impl=IMPL)
                    if zpad:
                        dw_ones = dw[:, ones_ch] if from_wgrad else None
                        dw = dw[:, :conv.weight.shape[1]].contiguous()
                    _acc(grads, conv.weight, dw)
                if has_bias:
                    _acc(grads, conv.bias, db)
                if from_wgrad:
                    _first_norm_grads(recs[0], conv, dw, dw_ones, grads)
            if from_wgrad:
                return None, False, None
            if last and not need_dx:
                return None, False, None
            if wp_bwd is None:      # forward ran on the CUDA-core first-layer kernel: pack for the data gradient now
                wpad = torch.nn.functional.pad(conv.weight.detach(), (0, 0) * (conv.weight.dim() - 2) + (0, zpad))
                _, wp_bwd, k = ops.conv_repack_weight(wpad, need_bwd=True)
            # the fused mask must have the layout of the (dense) dx the kernel writes
            fuse = bool(mask_in and not zpad and x_in.is_full())
            want, dot = _consumer_plan(nxt_consumer, fuse)
            # Epilogue statistics ride for free only under a long enough MMA phase (taps * C_out of this
            # dgrad); on shallow 2D layers the epilogue is the bottleneck and a separate HBM-speed pass
            # is cheaper (measured: profiles/r01_epilogue_probe.log)
            if want and dot is None and conv.weight.shape[0] * k[0] * k[1] * k[2] < FUSE_SUM_MIN_K:
                want = False
            if want and dot is not None:
                dot = dot.widened()
                if dot.c != x_in.c or dot.dims != x_in.dims:
                    want, dot = False, None
            fold = nxt_consumer[7] if (nxt_consumer is not None and nxt_consumer[0] == "gn") else None
            if fold is not None and not fold.done:
                # The norm's input was never written beyond its skip half (UpFold).  A norm without activation has
                # y = gamma * xhat + beta, and y is this conv's own input: correlate the gradient with y instead of
                # x and let the norm backward undo the affine map (ops.gn_bwd_fused_up, beta_if_y).
                if (want and dot is not None and FOLD_UPSAMPLE_BWD and up_fold is not None and idx == 1
                        and nxt_consumer[5] is None and not zpad and x_in.is_full() and nxt_consumer[1].is_full()):
                    dot, fold.dot_is_y = x_in, True
                else:
                    fold.materialise()
            if want:   # only sums are consumed unless a norm's (sum g, sum g*x) is asked for
                dy, sums = ops.conv_bwd_data(dy, wp_bwd, k, mask=x_in if fuse else None, impl=IMPL,
                                             stats="sum", dot=dot)
            else:
                dy = ops.conv_bwd_data(dy, wp_bwd, k, mask=x_in if fuse else None, impl=IMPL)
            out_premasked = fuse
            if zpad:  # gradient w.r.t. the padding channels is meaningless: expose the real slice
                dy = CL(dy.buf, dy.c - zpad, 0, zpad)
                if sums is not None:
                    sums = sums[:, :, :dy.c].contiguous()
        elif kind == "gn":
            _, x_in, gn, stats, gamma, y_relu, mask_in, fold = rec
            if y_relu is not None and not pm:
                dy, _ = ops.relu_bwd(dy, y_relu, False, out=dy)
                sums = None
            fused_up = (sums is not None and up_fold is not None and idx == 0 and FOLD_UPSAMPLE_BWD
                        and x_in.is_full() and up_fold[0] % 8 == 0 and (x_in.c - up_fold[0]) % 8 == 0)
            if fold is not None and not fold.done and not (fused_up and fold.dot_is_y):
                if fold.dot_is_y:
                    raise RuntimeError("statistics were taken against the norm's output but its folded backward "
                                       "is not available")      # cannot happen: same conditions on both sides
                fold.materialise()      # the passes below read the whole concat
            if fused_up:
                folded = fold is not None and not fold.done
                dskip, dlow, dgamma, dbeta = ops.gn_bwd_fused_up(
                    x_in, gn.num_groups, stats, gamma, dy, sums, up_fold[0], up_fold[1], mask_relu=mask_in,
                    x_low=fold.low if folded else None,
                    beta_if_y=gn.bias.detach().float().contiguous() if folded else None)
                dx = FoldedCatGrad(dskip, dlow)
            elif sums is not None:
                dx, dgamma, dbeta = ops.gn_bwd_fused(x_in, gn.num_groups, stats, gamma, dy, sums,
                                                     mask_relu=mask_in)
            else:
                dx, dgamma, dbeta = ops.gn_bwd(x_in, gn.num_groups, stats, gamma, dy, mask_relu=mask_in)
            sums = None
            _acc(grads, gn.weight, dgamma)
            _acc(grads, gn.bias, dbeta)
            dy = dx
            out_premasked = bool(mask_in)
        elif kind == "act":
            _, akind, param, y_act = rec
            if not (akind == ops.ACT_RELU and pm):
                dy = ops.act_bwd(dy, y_act, akind, param, out=dy)
            sums = None
            out_premasked = False
        elif kind == "drop":
            _, keep, chan, scale = rec
            dy = ops.scale_mask(dy, keep, chan, scale, out=dy)
            sums = None
            out_premasked = False
        elif kind == "bn":
            _, x_in, bn, mean, rstd, gamma, y_relu, was_training, mask_in = rec
            if not was_training or (y_relu is not None and not pm):
                sums = None
            # eval-mode BN (frozen statistics) is a per-channel affine map: no mean/variance terms
            dx, dgamma, dbeta = ops.bn_bwd(x_in, dy, None if pm else y_relu, mean, rstd, gamma,
                                           frozen=not was_training, mask_relu=mask_in, colstats=sums)
            sums = None
            _acc(grads, bn.weight, dgamma)
            _acc(grads, bn.bias, dbeta)
            dy = dx
            out_premasked = bool(mask_in)
    return dy, out_premasked, sums


def block_fwd(x: CL, specs: List[LayerSpec], training: bool, tape: Optional[list],
              out: Optional[CL] = None, in_relu: bool = False,
              in_sums: Optional[torch.Tensor] = None, want_sums: bool = False,
              up_src: Optional[UpFold] = None):
    """A DoubleConv / DoubleConvolution / unetConv2: consecutive layers, last one writes ``out``.
    Returns (y, y_is_relu_output, y_sums) — see single_conv_fwd for the sums."""
    flag, sums = in_relu, in_sums
    for j, spec in enumerate(specs):
        sub = [] if tape is not None else None
        is_last = j == len(specs) - 1
        want = want_sums if is_last else specs[j + 1].order[0] in "gb"
        x, flag, sums = single_conv_fwd(x, spec, training, sub, out=out if is_last else None,
                                        in_relu=flag, in_sums=sums, want_sums=want,
                                        up_src=up_src if j == 0 else None)
        if tape is not None:
            tape.append(sub)
    return x, flag, sums


def block_bwd(dy: CL, tape: list, nlayers: int, grads: Grads, need_dx: bool = True,
              premasked: bool = False, dy_sums: Optional[torch.Tensor] = None, consumer=None, up_fold=None):
    """Returns (dx, dx_premasked, dx_sums) — see single_conv_bwd (``up_fold`` applies to the first layer)."""
    pm, sums = premasked, dy_sums
    for j in range(nlayers):
        recs = tape.pop()
        inner = j < nlayers - 1
        nxt = tape[-1][-1] if inner else consumer   # the previous layer's last op receives our dx
        dy, pm, sums = single_conv_bwd(dy, recs, grads, need_dx=need_dx or inner, premasked=pm,
                                       dy_sums=sums, consumer=nxt, up_fold=None if inner else up_fold)
    return dy, pm, sums


# ===================================================================================== 2D U-Net
def unet2d_specs(block) -> List[LayerSpec]:
    """DoubleConvolution (model/unet2d/layers.py:103-133): conv+ReLU twice, bias, no norm."""
    return [LayerSpec("cr", block.first), LayerSpec("cr", block.second)]


def _crop_offsets(sh: int, sw: int, h: int, w: int) -> Tuple[int, int]:
    return int(round((sh - h) / 2.0)), int(round((sw - w) / 2.0))


def unet2d_forward(net, images: torch.Tensor, training: bool, tape: Optional[list]) -> CL:
    """model/unet2d/unet.py:95-128 up to (excluding) the 1x1 head; returns the 64-channel feature map.

    Concat order is [up-sampled, skip] (layers.py:188): the ConvTranspose epilogue scatters into
    channels [0, C) and the encoder's second conv writes channels [C, 2C) of the same buffer."""
    x = ops.from_nchw(images, pad=True)
    n_lvl = len(net.down_conv)
    skips: List[CL] = []
    relu = False          # is x a ReLU output?  (lets backward fold ReLU' into its neighbours)
    for i in range(n_lvl):
        c = net.down_conv[i].second.out_channels
        if x.h % 2 == 0 and x.w % 2 == 0 and _fits_after(x.h, x.w, n_lvl - i):
            cat = CL.empty(x.n, 1, x.h, x.w, 2 * c, x.device)
            dst = cat.slice(c, c)
        else:  # odd sizes need the centre crop of CropAndConcat: keep the skip in its own buffer
            dst = None
        x, relu, _ = block_fwd(x, unet2d_specs(net.down_conv[i]), training, tape, out=dst, in_relu=relu)
        skips.append(x)
        x = ops.maxpool_fwd(x, 1)
    x, relu, _ = block_fwd(x, unet2d_specs(net.middle_conv), training, tape, in_relu=relu)
    for i in range(n_lvl):
        skip = skips[n_lvl - 1 - i]
        up = net.up_sample[i].up
        wp_fwd, wp_bwd = ops.convt_repack_weight(up.weight)
        bias = None if up.bias is None else up.bias.detach().float().contiguous()
        c = up.out_channels
        if skip.ctot == 2 * c and skip.coff == c and (skip.h, skip.w) == (2 * x.h, 2 * x.w):
            cat = CL(skip.buf)
            crop = None
        else:
            cat = CL.empty(x.n, 1, 2 * x.h, 2 * x.w, c + skip.c, x.device)
            top, left = _crop_offsets(skip.h, skip.w, 2 * x.h, 2 * x.w)
            cropped = skip.dense()[:, :, top:top + 2 * x.h, left:left + 2 * x.w, :]
            cat.buf[..., c:] = cropped
            crop = (top, left)
        ops.convt_fwd(x, wp_fwd, bias, out=cat.slice(0, c), impl=IMPL)
        if tape is not None:
            tape.append(("up", x, up, wp_bwd, c, crop, skip, relu))
        # the concat mixes a ConvTranspose output with a ReLU output: not maskable as a whole
        x, relu, _ = block_fwd(cat, unet2d_specs(net.up_conv[i]), training, tape, in_relu=False)
    if tape is not None:
        tape.append(("skips", skips, relu))
    return x


def _fits_after(h: int, w: int, levels_left: int) -> bool:
    """True when every remaining pool halves exactly, so centre-crop is the identity."""
    for _ in range(levels_left):
        if h % 2 or w % 2:
            return False
        h //= 2
        w //= 2
    return True


def unet2d_backward(net, tape: list, dfeat: CL, grads: Grads, premasked: bool) -> None:
    n_lvl = len(net.down_conv)
    _, skips, _ = tape.pop()
    dskips: List[Optional[CL]] = [None] * n_lvl
    d, pm, d_sums = dfeat, premasked, None
    for i in range(n_lvl - 1, -1, -1):
        # the ConvTranspose bias gradient is the column sum of its half of dcat: have dgrad produce it
        want_up_bias = net.up_sample[i].up.bias is not None
        dcat, _, dcat_sums = block_bwd(d, tape, 2, grads, premasked=pm, dy_sums=d_sums,
                                       consumer=("sum",) if want_up_bias else None)
        _, x_in, up, wp_bwd, c, crop, skip, x_relu = tape.pop()
        d_up = dcat.slice(0, c)
        if crop is None:
            dskips[n_lvl - 1 - i] = dcat.slice(c, skip.c)
        else:
            full = torch.zeros_like(skip.dense(), memory_format=torch.contiguous_format)
            top, left = crop
            full[:, :, top:top + dcat.h, left:left + dcat.w, :] = dcat.dense()[..., c:]
            dskips[n_lvl - 1 - i] = CL(full.contiguous())
        if up.bias is not None:
            if dcat_sums is not None:
                _acc(grads, up.bias, dcat_sums[:, 0, :c].sum(0).float())
            else:
                with _OnSide(d_up.buf):
                    _acc(grads, up.bias, ops.bias_grad(d_up))
        with _OnSide(x_in.buf, d_up.buf):
            _acc(grads, up.weight, ops.convt_bwd_weight(x_in, d_up, impl=IMPL))
        fuse = bool(x_relu and x_in.is_full())
        # x_in is the ReLU output of the block below: its second conv's bias gradient is the column
        # sum of the masked gradient this kernel writes
        if fuse and 4 * c >= FUSE_SUM_MIN_K:
            d, d_sums = ops.convt_bwd_data(d_up, wp_bwd, mask=x_in, impl=IMPL, stats="sum")
        elif fuse:
            d, d_sums = ops.convt_bwd_data(d_up, wp_bwd, mask=x_in, impl=IMPL), None
        else:
            d, d_sums = ops.convt_bwd_data(d_up, wp_bwd, mask=None, impl=IMPL), None
        pm = fuse
    d, pm, _ = block_bwd(d, tape, 2, grads, premasked=pm, dy_sums=d_sums)
    for i in range(n_lvl - 1, -1, -1):
        # skips are ReLU outputs of DoubleConvolution: fold their ReLU backward into the pool backward
        d = ops.maxpool_bwd(skips[i], d, 1, add=dskips[i], mask_relu=True)
        d, pm, _ = block_bwd(d, tape, 2, grads, need_dx=i > 0, premasked=True)


# ===================================================================================== 3D family
def single_conv_spec(sc, order: str) -> LayerSpec:
    """SingleConv (model/unet3d/buildingblocks.py:116-159): children 'conv', 'groupnorm', 'batchnorm',
    'dropout' / 'dropout2d'."""
    drop = getattr(sc, "dropout", None) or getattr(sc, "dropout2d", None)
    return LayerSpec(order, sc.conv, getattr(sc, "groupnorm", None), getattr(sc, "batchnorm", None),
                     dropout_p=drop.p if drop is not None else 0.0)


def double_conv_specs(dc, order: str) -> List[LayerSpec]:
    return [single_conv_spec(dc.SingleConv1, order), single_conv_spec(dc.SingleConv2, order)]


def unet3d_forward(net, x_in: torch.Tensor, training: bool, tape: Optional[list]) -> CL:
    """model/unet3d/model.py:125-143 up to the 1x1 head. Concat order is (skip, up-sampled)
    (buildingblocks.py:547): skip channels [0, Cs), up-sampled [Cs, Cs + Cx)."""
    is3d = net.is3d
    order = net.layer_order
    x = ops.from_nchw(x_in, pad=True)
    n_enc = len(net.encoders)
    skips: List[CL] = []
    relu = False
    skip_relu: List[bool] = []
    skip_sums: List[Optional[torch.Tensor]] = []
    x_sums: Optional[torch.Tensor] = None
    norm_first = order[0] in "gb"   # blocks start with a norm: their producers hand over statistics
    for i, enc in enumerate(net.encoders):
        if enc.pooling is not None:
            avg = getattr(enc, "pool_type", "max") == "avg"
            if not avg and norm_first and order[0] == "g":
                # the statistics of the GroupNorm that opens this level ride in the pooling pass
                pooled, pooled_sums = ops.maxpool_fwd_sums(x, 2 if is3d else 1)
            else:
                pooled, pooled_sums = (ops.avgpool_fwd if avg else ops.maxpool_fwd)(x, 2 if is3d else 1), None
            if tape is not None:
                tape.append(("pool", x, relu, avg))
            x, x_sums = pooled, pooled_sums
        cs = enc.basic_module.SingleConv2.conv.out_channels
        dst = None
        if i < n_enc - 1:
            cx = _decoder_up_channels(net, n_enc - 2 - i)
            cat = CL.empty(x.n, x.d, x.h, x.w, cs + cx, x.device)
            dst = cat.slice(0, cs)
        x, relu, x_sums = block_fwd(x, double_conv_specs(enc.basic_module, order), training, tape,
                                    out=dst, in_relu=relu, in_sums=x_sums,
                                    want_sums=norm_first and len(net.decoders) > 0)
        skips.append(x)
        skip_relu.append(relu)
        skip_sums.append(x_sums)
    for j, dec in enumerate(net.decoders):
        skip = skips[n_enc - 2 - j]
        cat = CL(skip.buf)
        cx = cat.ctot - skip.c
        size = (skip.d, skip.h, skip.w)
        up_dst = cat.slice(skip.c, cx)
        ct = getattr(getattr(dec.upsampling, "upsample", None), "conv_transposed", None)
        cat_sums: Optional[torch.Tensor] = None
        up_src: Optional[UpFold] = None
        if ct is not None:
            wp_fwd, wp_bwd, kd = ops.deconv_repack_weight(ct.weight)
            planes = ops.deconv_fwd(x, wp_fwd, kd, impl=IMPL)
            ops.deconv_resize_fwd(planes, x.n, kd, size, out=up_dst)
            if tape is not None:
                tape.append(("deconv", x, ct, wp_bwd, kd, skip.c, cx))
            cat_relu = False      # the transposed conv has no activation
        else:
            if tape is not None:
                tape.append(("nearest", (x.d, x.h, x.w), skip.c, cx))
            cat_relu = relu and skip_relu[n_enc - 2 - j]   # both halves are ReLU outputs
            # nearest up-sampling by whole factors replicates every voxel f times: the statistics of
            # the concat buffer are those of its two producers, no extra pass
            s_skip = skip_sums[n_enc - 2 - j]
            exact = all(o % i == 0 and ((o // i) & (o // i - 1)) == 0
                        for o, i in zip(size, (x.d, x.h, x.w)))   # power-of-two factors map exactly
            if x_sums is not None and s_skip is not None and exact:
                f = (size[0] // x.d) * (size[1] // x.h) * (size[2] // x.w)
                cat_sums = torch.cat([s_skip, x_sums * float(f)], dim=2)
            # with those statistics and 1x/2x factors the GroupNorm in front of the decoder's first conv reads
            # the low-resolution tensor itself: the up-sampled half of the concat is never written
            if (FOLD_UPSAMPLE_FWD and cat_sums is not None and order[0] == "g" and x.is_full()
                    and skip.c % 8 == 0 and cx % 8 == 0
                    and all(o == i or o == 2 * i for o, i in zip(size, (x.d, x.h, x.w)))):
                up_src = UpFold(cat, skip.c, x)
            else:
                ops.upsample_nearest_fwd(x, size, out=up_dst)
        if ct is not None:
            cat_sums = None
        x, relu, x_sums = block_fwd(cat, double_conv_specs(dec.basic_module, order), training, tape,
                                    in_relu=cat_relu, in_sums=cat_sums,
                                    want_sums=norm_first and j < len(net.decoders) - 1, up_src=up_src)
    if tape is not None:
        tape.append(("skips", skips, relu))
    return x


def _decoder_up_channels(net, j: int) -> int:
    """Channels of the up-sampled tensor entering decoder j's concat."""
    dec = net.decoders[j]
    ct = getattr(getattr(dec.upsampling, "upsample", None), "conv_transposed", None)
    if ct is not None:
        return ct.out_channels
    c_in = dec.basic_module.SingleConv1.conv.in_channels
    skip_c = net.encoders[len(net.encoders) - 2 - j].basic_module.SingleConv2.conv.out_channels
    return c_in - skip_c


def unet3d_backward(net, tape: list, dfeat: CL, grads: Grads, premasked: bool) -> None:
    is3d = net.is3d
    n_enc = len(net.encoders)
    _, skips, _ = tape.pop()
    dskips: List[Optional[CL]] = [None] * n_enc
    d, pm = dfeat, premasked
    for j in range(len(net.decoders) - 1, -1, -1):
        up_rec = tape[-3]       # [..., up-sampling record, layer 1 records, layer 2 records]
        up_fold = None
        if up_rec[0] == "nearest":
            _, in_size, cs_, _cx = up_rec
            skip = skips[n_enc - 2 - j]
            if all(o == i or o == 2 * i for o, i in zip((skip.d, skip.h, skip.w), in_size)):
                up_fold = (cs_, (skip.n,) + tuple(in_size))
        dcat, cat_pm, _ = block_bwd(d, tape, 2, grads, premasked=pm, up_fold=up_fold)
        pm = False   # up-sampling backward sums children: the result is not a masked gradient
        rec = tape.pop()
        if isinstance(dcat, FoldedCatGrad):        # the GroupNorm backward already reduced the up-sampled half
            d, pm = dcat.low, cat_pm
            dskips[n_enc - 2 - j] = dcat.skip
            continue
        if rec[0] == "deconv":
            _, x_in, ct, wp_bwd, kd, cs, cx = rec
            d_planes = ops.deconv_resize_bwd(dcat.slice(cs, cx), x_in.dims, kd)
            with _OnSide(x_in.buf, d_planes.buf):
                _acc(grads, ct.weight, ops.deconv_bwd_weight(x_in, d_planes, tuple(ct.weight.shape), kd,
                                                            impl=IMPL))
            d = ops.deconv_bwd_data(d_planes, wp_bwd, x_in.dims, kd, impl=IMPL)
        else:
            _, in_size, cs, cx = rec
            d = ops.upsample_nearest_bwd(dcat.slice(cs, cx), in_size)
            # all children of a source voxel share its value, so a gradient masked by (up > 0) sums to
            # a gradient masked by (source > 0): the ReLU backward of the producer is already in it
            pm = cat_pm
        dskips[n_enc - 2 - j] = dcat.slice(0, cs)
    for i in range(n_enc - 1, -1, -1):
        d, _, _ = block_bwd(d, tape, 2, grads, need_dx=i > 0, premasked=pm)
        if i > 0:
            _, pre, pre_relu, avg = tape.pop()
            # the skip half of the concat gradient joins the gradient routed back through the pool;
            # when the pooled tensor is a ReLU output its ReLU backward is folded in as well
            d = (ops.avgpool_bwd if avg else ops.maxpool_bwd)(pre, d, 2 if is3d else 1, add=dskips[i - 1],
                                                              mask_relu=pre_relu)
            pm = pre_relu


# ===================================================================================== autograd glue
class _NetFunction(torch.autograd.Function):
    """One autograd node for a whole encoder-decoder + 1x1 head; returns fp32 logits (N, C, ...)."""

    @staticmethod
    def forward(ctx, net, kind: str, images: torch.Tensor, *params):
        training = net.training
        need_grad = any(ctx.needs_input_grad)
        tape: Optional[list] = [] if need_grad else None
        with ops.on_device(images), torch.no_grad(), prepacked(net, need_grad):
            if kind == "unet2d":
                feat = unet2d_forward(net, images, training, tape)
            elif kind == "resunet3d":
                from . import resnet_exec
                feat = resnet_exec.forward(net, images, training, tape)
            else:
                feat = unet3d_forward(net, images, training, tape)
            head = net.final_conv
            bias = None if head.bias is None else head.bias.detach().float().contiguous()
            logits = ops.head_fwd(feat, head.weight, bias, images.dim() == 5)
        ctx.feat_relu = bool(tape[-1][2]) if need_grad else False
        ctx.net, ctx.kind, ctx.tape, ctx.feat = net, kind, tape, feat if need_grad else None
        ctx.params = params
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        net, tape = ctx.net, ctx.tape
        if tape is None:
            raise RuntimeError("backward through a forward that ran without grad")
        global _REDUCER
        grads: Grads = {}
        head = net.final_conv
        _REDUCER = getattr(net, "_b200_reducer", None)
        # measured (profiles/r01_wgrad_stream_ab.txt): -3 % on the 3D nets, whose backward carries ~5 ms of
        # HBM-bound norm / pool / up-sampling kernels to overlap; nothing on the norm-free 2D U-Net
        dev_ctx = ops.on_device(dlogits)
        dev_ctx.__enter__()
        if ctx.kind != "unet2d":
            _side_begin(dlogits.device)
        try:
            dfeat, dw, db = ops.head_bwd(ctx.feat, head.weight, dlogits, mask_relu=ctx.feat_relu)
            _acc(grads, head.weight, dw)
            if head.bias is not None:
                _acc(grads, head.bias, db)
            if ctx.kind == "unet2d":
                unet2d_backward(net, tape, dfeat, grads, ctx.feat_relu)
            elif ctx.kind == "resunet3d":
                from . import resnet_exec
                resnet_exec.backward(net, tape, dfeat, grads, ctx.feat_relu)
            else:
                unet3d_backward(net, tape, dfeat, grads, ctx.feat_relu)
            _side_end(grads)
            if _REDUCER is not None:
                grads.update(_REDUCER.finish())     # averaged gradients in fresh memory (never the bucket views)
        finally:
            _REDUCER = None
            _side_end({})
            dev_ctx.__exit__(None, None, None)
        ctx.tape = ctx.feat = None
        out = tuple(grads.get(p) if p.requires_grad else None for p in ctx.params)
        return (None, None, None) + out


def run_network(net, kind: str, images: torch.Tensor) -> torch.Tensor:
    params = tuple(net.parameters())
    return _NetFunction.apply(net, kind, images, *params)


class _BlockFunction(torch.autograd.Function):
    """Stand-alone use of a block module (DoubleConvolution, DoubleConv, unetConv2, SingleConv) on an
    fp32 NCHW tensor: same kernels, with the layout conversion at both ends."""

    @staticmethod
    def forward(ctx, mod, specs, x: torch.Tensor, *params):
        need_grad = any(ctx.needs_input_grad)
        tape = [] if need_grad else None
        with ops.on_device(x), torch.no_grad():
            y, _, _ = block_fwd(ops.from_nchw(x), specs, mod.training, tape)
            out = ops.to_nchw(y, x.dim() == 5)
        ctx.tape, ctx.n, ctx.params, ctx.x_grad = tape, len(specs), params, ctx.needs_input_grad[2]
        return out

    @staticmethod
    def backward(ctx, dout):
        grads: Grads = {}
        with ops.on_device(dout):
            dy = ops.from_nchw(dout)
            dx, _, _ = block_bwd(dy, ctx.tape, ctx.n, grads, need_dx=ctx.x_grad)
            gx = ops.to_nchw(dx, dout.dim() == 5) if ctx.x_grad else None
        return (None, None, gx) + tuple(grads.get(p) if p.requires_grad else None for p in ctx.params)


def run_block(mod, specs: List[LayerSpec], x: torch.Tensor) -> torch.Tensor:
    return _BlockFunction.apply(mod, specs, x, *tuple(mod.parameters()))


class _StandaloneFunction(torch.autograd.Function):
    """Stand-alone use of a module that is normally run by a whole-network executor (Encoder pooling,
    Decoder, ResNetBlock, SE layers): ``run(inputs, want_tape) -> (out fp32 NC[D]HW, bwd)`` launches the
    kernels; ``bwd(dout) -> (input grads, {param: grad})`` replays its tape."""

    @staticmethod
    def forward(ctx, run, n_in: int, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        need_grad = any(ctx.needs_input_grad)
        with ops.on_device(inputs[0]), torch.no_grad():
            out, bwd = run(inputs, need_grad)
        ctx.bwd, ctx.params, ctx.n_in = bwd, params, n_in
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.bwd is None:
            raise RuntimeError("backward through a forward that ran without grad")
        with ops.on_device(dout), torch.no_grad():
            dins, grads = ctx.bwd(dout.contiguous().float())
        ctx.bwd = None
        dins = tuple(d if need else None for d, need in zip(dins, ctx.needs_input_grad[2:2 + ctx.n_in]))
        return (None, None) + dins + tuple(grads.get(p) if p.requires_grad else None for p in ctx.params)


def run_standalone(run, inputs, params) -> torch.Tensor:
    return _StandaloneFunction.apply(run, len(inputs), *inputs, *params)


def run_pool(x: torch.Tensor, is3d: bool, avg: bool) -> torch.Tensor:
    """2x max / average pooling of an fp32 NC[D]HW tensor through the path's kernels (Encoder.forward)."""
    pd = 2 if is3d else 1

    def run(inputs, want):
        xin = ops.from_nchw(inputs[0])
        y = (ops.avgpool_fwd if avg else ops.maxpool_fwd)(xin, pd)

        def bwd(dout):
            d = (ops.avgpool_bwd if avg else ops.maxpool_bwd)(xin, ops.from_nchw(dout), pd)
            return (ops.to_nchw(d, is3d),), {}
        return ops.to_nchw(y, is3d), (bwd if want else None)
    return run_standalone(run, (x,), ())
