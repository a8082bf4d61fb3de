"""CUDA-graph replay of a fixed-shape inference forward.

The reference serves ``model(input_batch)`` per request (api/api.py:64) and predicts patch by patch
(model/unet3d/predictor.py:128); at batch 1 a U-Net forward is ~45 (2D) / ~60 (3D) short kernels and the
Python + ctypes launch path, not the GPU, sets the latency.  Every C-ABI call enqueues on the caller's
stream without allocating or synchronising, so a whole forward can be captured once per
input shape and replayed with one launch.  Weights are re-packed from the fp32 masters inside the graph, so
``load_state_dict`` / optimizer updates between replays are honoured.
"""
from __future__ import annotations

import os

import torch


class GraphedInference:
    """``g = GraphedInference(model, example); logits = g(x)`` — x must have example's shape and dtype.
    The returned tensor is a static buffer that the next call overwrites (clone it to keep it)."""

    def __init__(self, module: torch.nn.Module, example: torch.Tensor, warmup: int = 2):
        if not example.is_cuda:
            raise RuntimeError("GraphedInference needs a CUDA example input (no CPU fallback)")
        self.module = module.eval()
        self.static_in = example.detach().clone()
        side = torch.cuda.Stream(device=example.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):       # first calls configure kernel attributes: keep them out
                self.module(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(example.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self.module(self.static_in)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"captured for {tuple(self.static_in.shape)} {self.static_in.dtype}, "
                             f"got {tuple(x.shape)} {x.dtype}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """One whole training step — forward, loss, backward, optimizer update — captured once as a CUDA graph and
    replayed with a single launch:

        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=True)
        step = GraphedTrainStep(lambda x, y: criterion(model(x), y), opt, (x0, y0))
        for x, y in loader:
            loss = step(x, y)          # static tensor, overwritten by the next call

    A cfg3 step is ~120 library calls plus the optimizer's and allocator's own kernels — ~20 ms of Python + ctypes
    enqueue work against ~24 ms on the device: the eager loop is GPU-bound only while the host keeps up (measured:
    one eager bench run in six lost 15-30 % to a slow host moment, none of the graphed ones).  Every
    C-ABI call enqueues on the caller's stream without allocating on the host side of the device or synchronising
    (SURVEY.md section 8b), the weight-gradient side stream forks and joins with events, accumulators come from
    in-graph memsets, and the filter packs are rebuilt from the fp32 masters inside the graph — so the whole step
    captures as is and replays honour the optimizer's updates.  Inputs must keep the example's shapes and dtypes;
    the optimizer must be capturable (``capturable=True`` for Adam / AdamW).  Under ``DataParallelUNet`` the
    bucketed NCCL all-reduces of the gradient reducer are captured with the step (its communication stream joins
    the capture through the events it already waits on); every rank must build and replay its graph in step, and
    drop it (``del step``) before ``destroy_process_group()`` — a live graph that holds NCCL work keeps the
    communicator from shutting down."""

    def __init__(self, loss_fn, optimizer: torch.optim.Optimizer, example_inputs, warmup: int = 3):
        example_inputs = tuple(example_inputs)
        if not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedTrainStep needs CUDA example inputs (no CPU fallback)")
        self.loss_fn, self.optimizer = loss_fn, optimizer
        self.static_in = tuple(t.detach().clone() for t in example_inputs)
        device = example_inputs[0].device
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):       # kernel attributes, filter-pack plan, optimizer state
                optimizer.zero_grad(set_to_none=True)
                loss_fn(*self.static_in).backward()
                optimizer.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(device)
        from . import _lib
        optimizer.zero_grad(set_to_none=True)
        # the warm-up activations were cached for the side stream; hand them back before the capture takes the same
        # amount again for its private pool (UNet 3+ at 512^2 x 8: ~45 GB per copy)
        torch.cuda.empty_cache()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count
        # capture on the warm-up stream: the parameters' AccumulateGrad nodes were created there (autograd warns when
        # a later backward reaches them from another stream — eager steps after this object do, harmlessly)
        try:
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        except AttributeError:
            pass
        from . import ops
        if os.environ.get("B200_GRAPH_ZERO_POOL", "1") != "0":      # A/B knob
            ops._ZEROS.begin_capture()      # small accumulators: a few in-graph memsets instead of ~100
        try:
            with torch.cuda.graph(self.graph, stream=side):
                self.static_loss = loss_fn(*self.static_in)
                self.static_loss.backward()
                optimizer.step()
        finally:
            ops._ZEROS.end_capture()
        self.launches_per_step = _lib.launch_count - before     # C-ABI kernel launches inside one replay

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        if len(inputs) != len(self.static_in):
            raise ValueError(f"captured for {len(self.static_in)} inputs, got {len(inputs)}")
        for dst, src in zip(self.static_in, inputs):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError(f"captured for {tuple(dst.shape)} {dst.dtype}, got {tuple(src.shape)} {src.dtype}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_loss
