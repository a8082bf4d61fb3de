"""Data-parallel training: one process per GPU, gradients averaged with bucketed NCCL all-reduces that
overlap the rest of the backward pass.

The reference's only multi-GPU mechanism is single-process ``nn.DataParallel``
(model/unet3d/trainer.py:23-24): per-step weight broadcast, full-resolution logits gathered on GPU 0,
gradients reduced onto GPU 0.  Here every rank owns a replica and a batch shard; the whole-network
backward (engine._NetFunction) hands each parameter gradient to :class:`GradReducer` the moment it is
final — decoder and head gradients first — which packs it into a flat fp32 bucket and, once a bucket is
full, enqueues ``all_reduce(sum) / world`` on a communication stream ordered after the producing
kernels by an event.  The compute stream only waits for the communication stream at the end of
backward, so all but the last bucket's transfer hides behind dgrad/wgrad kernels.  BatchNorm statistics
stay per replica, as under DataParallel.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.distributed as dist
from torch import nn


class GradReducer:
    def __init__(self, params, bucket_mb: float = 32.0, group=None):
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.bucket_elems = max(1, int(bucket_mb * 1024 * 1024 / 4))
        self.cuda = dev.type == "cuda"
        self.comm_stream = torch.cuda.Stream(device=dev) if self.cuda else None
        self._reset()
        self.launched_buckets = 0

    def _reset(self):
        self.offset = 0
        self.bucket_start = 0
        self.handles = []
        self.slots: Dict[nn.Parameter, torch.Tensor] = {}
        self._writers = set()      # streams that wrote slots of the bucket being filled

    # ---- called from engine backward, in the order gradients become final
    def on_grad(self, p: nn.Parameter, g: torch.Tensor) -> torch.Tensor:
        """Packs a final gradient into its bucket slot (on the CURRENT stream — the executor calls this from
        its main stream and from its weight-gradient side stream) and returns the slot.  The slot is a view
        of the persistent bucket memory: it is only valid until the next backward; ``finish()`` hands out the
        tensors that may become ``p.grad``."""
        n = p.numel()
        view = self.flat[self.offset:self.offset + n]
        if self.cuda:
            self._writers.add(torch.cuda.current_stream())
        if self.world > 1:
            torch.div(g.reshape(-1), float(self.world), out=view)
        else:
            view.copy_(g.reshape(-1))
        self.slots[p] = view.view(p.shape)
        self.offset += n
        if self.offset - self.bucket_start >= self.bucket_elems:
            self._launch()
        return self.slots[p]

    def _launch(self):
        if self.offset == self.bucket_start:
            return
        chunk = self.flat[self.bucket_start:self.offset]
        self.bucket_start = self.offset
        if self.world <= 1:
            return
        self.launched_buckets += 1
        if self.cuda:
            # every stream that wrote a slot of this bucket (main AND side stream) must be ahead of the
            # all-reduce, not only the one that happened to cross the bucket boundary
            for st in self._writers | {torch.cuda.current_stream()}:
                ev = torch.cuda.Event()
                ev.record(st)
                self.comm_stream.wait_event(ev)
            self._writers = set()
            with torch.cuda.stream(self.comm_stream):
                self.handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group,
                                                    async_op=True))
        else:
            self.handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group,
                                                async_op=True))

    def finish(self) -> Dict[nn.Parameter, torch.Tensor]:
        """Flush the last bucket, order the compute stream after every all-reduce and return the averaged
        gradients as views of ONE freshly allocated tensor: autograd keeps what backward returns as
        ``p.grad`` without cloning, so handing out views of the persistent bucket memory would let the next
        backward overwrite ``p.grad`` in place (gradient accumulation / ``zero_grad(set_to_none=False)``)."""
        self._launch()
        for h in self.handles:
            if self.cuda:
                with torch.cuda.stream(self.comm_stream):
                    h.wait()
            else:
                h.wait()
        if self.cuda:
            cur = torch.cuda.current_stream()
            if self.world > 1:
                cur.wait_stream(self.comm_stream)
            for st in self._writers:          # world == 1: no all-reduce ordered the writers before us
                if st != cur:
                    cur.wait_stream(st)
        fresh = self.flat[:self.offset].clone()
        out, off = {}, 0
        for p, slot in self.slots.items():    # insertion order == packing order
            n = slot.numel()
            out[p] = fresh[off:off + n].view(slot.shape)
            off += n
        self._reset()
        return out


class DataParallelUNet(nn.Module):
    """Wraps a network of this package for one-process-per-GPU data parallelism.  Parameters are
    broadcast from rank 0 at construction; every backward averages gradients across ranks through a
    :class:`GradReducer` that the executor feeds while it is still computing."""

    def __init__(self, module: nn.Module, bucket_mb: float = 32.0, group=None):
        super().__init__()
        self.module = module
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=group)
        self.reducer = GradReducer(module.parameters(), bucket_mb=bucket_mb, group=group)
        for m in module.modules():
            # whole-network executors: the U-Nets (final_conv) and the UNet 3+ family (_b200_executor)
            if (hasattr(m, "final_conv") or getattr(m, "_b200_executor", False)) and hasattr(m, "forward"):
                m._b200_reducer = self.reducer

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def allreduce_mean_(t: torch.Tensor, group=None) -> torch.Tensor:
    """Tiny helper for loss / metric partial sums (one small all-reduce per step)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t /= dist.get_world_size(group)
    return t
