#!/usr/bin/env python
"""Headline benchmark: U-Net training throughput on B200 (imgs/s for 2D, vols/s for 3D).

    python bench.py --gpus N --steps K --warmup W [--workload cfg2|cfg3|cfg3p|cfg4|cfg4p|cfg1|u3p|u3pds|res3d|res3dse|cfg5] [--impl reference]

A step = forward + loss + backward + optimizer step on one per-GPU batch of synthetic data (random-init
weights of the named architecture).  One JSON line on rank 0 (see the keys at the bottom).  For N > 1
the driver launches this file under torch.distributed.run; ranks shard by batch (weak scaling) and the
gradient all-reduce overlaps backward (mdeical_image_segmentation_b200/parallel.py).

Default workload = cfg3 (the north-star configuration: UNet3D(1,1) gcr + nearest, 128^3, batch 2 per GPU);
the same line carries short device-resident measurements of cfg2 and cfg3p under "extra".

The timed steps are replays of ONE CUDA graph per rank holding the whole step (GraphedTrainStep, the package's public
API for a fixed-shape training loop; the NCCL all-reduces are inside the graph at N > 1) -- "step_launch" in the JSON
line says so, --no-graph times the eager loop.  The roofline / breakdown steps are eager with CUDA events around every
library call; "roofline" names the SM clock of both regimes and is held against the cuBLAS peak of the regime its
kernels were timed in.

`--impl reference` times the reference's own CPU implementation of the same path on all host threads: the
UNMODIFIED reference modules from oracle/_ref (a copy of the path's files made by oracle/build_ref.py in the
build container; git-ignored, travels with the snapshot) -- "kind": "reference" -- or, when oracle/_ref is
absent, the oracle port (oracle/unet.py, pinned to the reference by tests/golden) -- "kind": "port".  A step
of that arm is a bounded sample of the workload (the per-GPU batch of 64^3 crops / 256^2 crops, throughput
scaled by voxel count) so that --steps / --warmup are honoured as given.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# conv + transposed-conv FLOPs per sample, train = 3 x fwd - fwd(first conv)  
WORKLOADS = {
    "cfg1": dict(kind="2d", cin=1, cout=1, size=(256, 256), batch=4, train_tf=0.2885, unit="imgs/s",
                 desc="UNet(1,1) 256x256 binary seg, batch 4, BCE + IoU/Dice metrics"),
    "cfg2": dict(kind="2d", cin=3, cout=4, size=(512, 512), batch=64, train_tf=1.1554, unit="imgs/s",
                 desc="UNet(3,4) 512x512 4-class, batch 64 per GPU, CrossEntropy + AdamW"),
    "cfg3": dict(kind="3d", cin=1, cout=1, size=(128, 128, 128), batch=2, train_tf=11.358, unit="vols/s",
                 order="gcr", upsample="default",
                 desc="UNet3D(1,1) gcr+nearest 128^3, batch 2 per GPU, BCEDice + Adam"),
    "cfg3p": dict(kind="3d", cin=1, cout=1, size=(128, 128, 128), batch=2, train_tf=9.532, unit="vols/s",
                  order="cbr", upsample="deconv",
                  desc="UNet3D(1,1) cbr+deconv 128^3, batch 2 per GPU, BCEDice + Adam"),
    "cfg4": dict(kind="3d", cin=4, cout=3, size=(160, 192, 160), batch=1, train_tf=26.675, unit="vols/s",
                 order="gcr", upsample="default", loss="dice",
                 desc="UNet3D(4,3) gcr+nearest 160x192x160 (BraTS-like), batch 1 per GPU, fused Dice + Adam"),
    "cfg4p": dict(kind="3d", cin=4, cout=3, size=(160, 192, 160), batch=1, train_tf=22.394, unit="vols/s",
                  order="cbr", upsample="deconv", loss="dice",
                  desc="UNet3D(4,3) cbr+deconv 160x192x160 (BraTS-like), batch 1 per GPU, fused Dice + Adam"),
    # ResidualUNet3D(1,1), 5 levels 64..1024 (model/unet3d/model.py:197-238)
    "res3d": dict(kind="3d", cin=1, cout=1, size=(128, 128, 128), batch=2, train_tf=11.2885, unit="vols/s",
                  order="gcr", upsample="default", residual=True,
                  desc="ResidualUNet3D(1,1) gcr, deconv + summation joining, 128^3, batch 2 per GPU, BCEDice + Adam"),
    "res3dse": dict(kind="3d", cin=1, cout=1, size=(128, 128, 128), batch=2, train_tf=11.2885, unit="vols/s",
                    order="gcr", upsample="default", residual="se",
                    desc="ResidualUNetSE3D(1,1) gcr (scSE gate after every block), 128^3, batch 2 per GPU, BCEDice + Adam"),
    # the default model of the reference's train.py (train.py:101-105)
    "u3p": dict(kind="2d", cin=3, cout=1, size=(512, 512), batch=8, train_tf=4.7428, unit="imgs/s",
                unet_type="UNet_3Plus",
                desc="UNet_3Plus(3,1) 512x512 binary seg, batch 8 per GPU, SegmentationLoss + AdamW"),
    "u3pds": dict(kind="2d", cin=3, cout=1, size=(512, 512), batch=8, train_tf=4.7443, unit="imgs/s",
                  unet_type="UNet_3Plus_DeepSup",
                  desc="UNet_3Plus_DeepSup(3,1) 512x512 binary seg, batch 8 per GPU, SegmentationLoss over 5 heads + AdamW"),
}

# cfg5 (BASELINE.json configs[4]): inference-only sweep, eval() + no_grad, as api/api.py:64 would call it.
# (kind, batch, size, fwd TFLOP per sample)
SWEEP = [("2d", 1, (256, 256), 0.0962), ("2d", 16, (256, 256), 0.0962), ("2d", 64, (256, 256), 0.0962),
         ("2d", 256, (256, 256), 0.0962), ("2d", 1, (512, 512), 0.385), ("2d", 16, (512, 512), 0.385),
         ("2d", 64, (512, 512), 0.385), ("2d", 1, (1024, 1024), 1.539), ("2d", 16, (1024, 1024), 1.539),
         ("3d", 1, (96, 96, 96), 1.598), ("3d", 4, (96, 96, 96), 1.598), ("3d", 1, (128, 128, 128), 3.787),
         ("3d", 2, (128, 128, 128), 3.787), ("3d", 8, (128, 128, 128), 3.787), ("3d", 1, (160, 160, 160), 7.40),
         ("3d", 1, (192, 192, 192), 12.78)]


_JSON_FD = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner comes
    through printf from inside init) is sent to stderr by pointing fd 1 at fd 2; the JSON line goes to a
    duplicate of the original fd 1."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------- data
def make_batch(w, batch, seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    if w["kind"] == "2d":
        x = torch.randn((batch, w["cin"]) + w["size"], generator=g)
        if w["cout"] > 1:
            y = torch.randint(0, w["cout"], (batch,) + w["size"], generator=g)
        else:
            y = (torch.rand((batch, 1) + w["size"], generator=g) > 0.7).float()
    elif w["cout"] > 1:   # one-hot float target of the input's shape (model/unet3d/losses.py:20)
        x = torch.randn((batch, w["cin"]) + w["size"], generator=g)
        cls = torch.randint(0, w["cout"], (batch,) + w["size"], generator=g)
        y = torch.nn.functional.one_hot(cls, w["cout"]).movedim(-1, 1).float().contiguous()
    else:
        x = torch.randn((batch, w["cin"]) + w["size"], generator=g)
        y = (torch.rand((batch, w["cout"]) + w["size"], generator=g) > 0.7).float()
    if pin:
        x, y = x.pin_memory(), y.pin_memory()
    return x.to(device), y.to(device)


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index, period=0.1):
        self.period = period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80, "sync_boost": 0x10,
                 "applications_clocks": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------------------- B200 arm
def build_model(w, device, capturable=False):
    """capturable: optimizer state on the device, so that optimizer.step() can sit inside a CUDA graph."""
    import mdeical_image_segmentation_b200 as B
    torch.manual_seed(0)
    if w["kind"] == "2d":
        model = B.UNetModel(B.UNetConfig(w["cin"], w["cout"], w.get("unet_type", "UNet"))).to(device).train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-3, fused=True, capturable=capturable)

        def step_fn(x, y):
            return model(images=x, labels=y).loss
    else:
        if w.get("residual"):
            cls = B.ResidualUNetSE3D if w["residual"] == "se" else B.ResidualUNet3D
            net = cls(w["cin"], w["cout"], layer_order=w["order"]).to(device).train()
        else:
            net = B.UNet3D(w["cin"], w["cout"], layer_order=w["order"], upsample=w["upsample"]).to(device).train()
        crit = B.DiceLoss(normalization="sigmoid") if w.get("loss") == "dice" else B.BCEDiceLoss(1.0, 1.0)
        model = net
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True, capturable=capturable)

        def step_fn(x, y):
            return crit(net(x), y)
    return model, opt, step_fn


def conv_flops(name, args, real_cin=None):
    """Algorithmic FLOPs of one C-ABI conv-family call, from its descriptor.  The narrow network input is
    stored zero-padded to 16 channels (and its 3D weight gradient runs over an im2col with K' = 32): those
    calls count the REAL input channels, 2 * positions * taps * C_in * C_out."""
    d = args[0]._obj if hasattr(args[0], "_obj") else None
    if d is None:
        return 0.0
    if name.startswith("b200_convt2x2"):
        return 2.0 * d.n * d.h * d.w * d.c_in * d.c_out * 4
    if name.startswith("b200_conv_"):
        c_in, c_out, taps = d.c_in, d.c_out, d.kd * d.kh * d.kw
        if real_cin is not None and real_cin < 16:
            full_taps = 27 if d.d > 1 else 9
            kp = (full_taps * real_cin + 15) // 16 * 16
            if c_in == 16 and taps == full_taps:
                c_in = real_cin                 # first layer on the padded input (fwd / dgrad / wgrad descriptors)
            elif taps == 1 and c_in == kp:
                c_in, taps = real_cin, full_taps   # im2col weight gradient: K' = taps * C_in rounded up
        return 2.0 * d.n * d.d * d.h * d.w * c_in * c_out * taps
    if name.startswith("b200_deconv3s2"):
        return 2.0 * d.n * d.d * d.h * d.w * d.c_in * d.c_out * (27 if d.d > 1 else 9)
    return 0.0


class KernelProfiler:
    """Per-call CUDA-event timing on the launching stream for one instrumented step."""

    def __init__(self, real_cin=None, steps=1):
        self.records = []
        self.real_cin = real_cin
        self.steps = steps          # instrumented steps recorded: families are averaged over them

    @contextlib.contextmanager
    def hook(self, name, args):
        from mdeical_image_segmentation_b200 import _lib
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        yield
        e1.record()
        impl = _lib.lib().b200_last_impl() if ("conv" in name and "repack" not in name) else 0
        self.records.append((name, impl, conv_flops(name, args, self.real_cin), e0, e1))

    def summarize(self):
        torch.cuda.synchronize()
        fam = {}
        for name, impl, fl, e0, e1 in self.records:
            ms = e0.elapsed_time(e1)
            if "repack" in name:
                key = "repack"
            elif name in ("b200_conv_fwd", "b200_conv_bwd_data", "b200_convt2x2_fwd", "b200_convt2x2_bwd_data",
                          "b200_conv_fwd_stats", "b200_conv_bwd_data_stats", "b200_convt2x2_bwd_data_stats"):
                key = "tc_gemm_kernel" if impl == 1 else "direct_gemm_kernel"
            elif name in ("b200_conv_bwd_weight", "b200_conv_bwd_weight_bias", "b200_convt2x2_bwd_weight"):
                key = "tc_wgrad_kernel" if impl == 1 else "direct_wgrad_kernel"
            elif name in ("b200_deconv3s2_fwd", "b200_deconv3s2_bwd_data"):
                key = "tc_gemm_kernel(deconv)" if impl == 1 else "direct_gemm_kernel(deconv)"
            elif name == "b200_deconv3s2_bwd_weight":
                key = "tc_wgrad_kernel(deconv)" if impl == 1 else "direct_wgrad_kernel(deconv)"
            else:
                key = name.replace("b200_", "")
            f = fam.setdefault(key, dict(ms=0.0, flops=0.0, launches=0))
            f["ms"] += ms / self.steps
            f["flops"] += fl / self.steps
            f["launches"] += 1.0 / self.steps
        for f in fam.values():
            f["launches"] = int(round(f["launches"]))
        return fam


REFERENCE_SAMPLE = {
    "3d": "the --impl reference arm and cpu_baseline time the per-GPU batch of 64^3 crops per step on the host "
          "cores, throughput scaled by voxel count to the workload's volume",
    "2d": "the --impl reference arm and cpu_baseline time one 512x512 image per step on the host cores "
          "(UNet 3+: one 256x256 crop, scaled by pixel count)",
}


def traffic_for(workload, family):
    """DRAM bytes per launch of a kernel family from the newest committed ncu pass (STATIC: measured once
    under ncu by tools/ncu_summary.py, not in this run)."""
    for name in ("r02_dram_traffic.json", "r01_dram_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            v = json.load(open(path)).get(workload, {}).get(family, {}).get("dram_bytes_per_launch")
            if v is not None:
                return round(v), "profiles/" + name
    return None, None


def measure(args, name, w, steps, warmup, device, world, rank, full):
    """Times one workload on this rank's GPU.  full: also e2e, instrumented roofline step, launch count."""
    import torch.distributed as dist
    from mdeical_image_segmentation_b200 import _lib
    from mdeical_image_segmentation_b200 import engine as _engine
    from mdeical_image_segmentation_b200.parallel import DataParallelUNet

    batch = (args.batch if full else 0) or w["batch"]
    want_graph = args.graph
    model, opt, step_fn = build_model(w, device, capturable=want_graph)
    if world > 1:
        DataParallelUNet(model)
    x, y = make_batch(w, batch, 1234 + rank, device)

    def eager_step(xb, yb):
        opt.zero_grad(set_to_none=True)
        loss = step_fn(xb, yb)
        loss.backward()
        opt.step()
        return loss

    # The whole step (forward, loss, backward with its bucketed gradient all-reduces at N > 1, optimizer) is captured
    # once and replayed with a single launch per step — GraphedTrainStep, the package's public API for a fixed-shape
    # training loop.  The eager loop needs ~20 ms of host enqueue work per 24 ms cfg3 step, so any slow host moment
    # starves the GPU (and, data-parallel, every other rank with it).
    gstep = None
    if want_graph:
        import mdeical_image_segmentation_b200 as B
        try:
            gstep = B.GraphedTrainStep(step_fn, opt, (x, y), warmup=max(warmup, 3))
            x, y = gstep.static_in
        except Exception as e:   # noqa: BLE001 — a workload that cannot be captured is timed eagerly, and says so
            print(f"[bench] {name}: CUDA-graph capture failed ({type(e).__name__}: {str(e)[:200]}); eager steps",
                  file=sys.stderr)
            gstep = None
            torch.cuda.synchronize()
    if want_graph and world > 1:
        # every rank replays a graph or none does: the all-reduces of an eager rank would not pair with a peer's replay
        ok = torch.tensor([1.0 if gstep is not None else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1.0 and gstep is not None:
            gstep = None
            x, y = x.clone(), y.clone()
            release_device_memory()
    one_step = gstep if gstep is not None else eager_step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(warmup):
        one_step(x, y)
    barrier()
    _lib.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clocks:
        e0.record()
        for _ in range(steps):
            loss = one_step(x, y)
        e1.record()
        barrier()
    launches = _lib.launch_count if gstep is None else steps * gstep.launches_per_step
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item() / steps
    pk = peaks()
    step_tf = w["train_tf"] * batch
    res = {"value": world * batch / (ms_step / 1e3), "ms_step": ms_step, "batch": batch, "launches": launches,
           "step_launch": "one CUDA graph replay per step (GraphedTrainStep)" if gstep is not None else "eager",
           "clocks": clocks.summary(), "final_loss": float(loss.item()),
           "step": {"algorithmic_tflop_per_step": round(step_tf, 2),
                    "step_tflops": round(step_tf / (ms_step / 1e3), 1),
                    "frac_of_sustained_peak": round(step_tf / (ms_step / 1e3) / pk["tf_sustained"], 4)}}
    if not full:
        return res

    # ---- end to end: pinned host batch -> H2D every step (prefetched on a copy stream), loss -> host
    res["e2e"] = None
    if not args.no_e2e:
        host = [make_batch(w, batch, 99 + rank + i, "cpu", pin=True) for i in range(2)]
        copy_stream = torch.cuda.Stream(device=device)
        h2d = sum(t.numel() * t.element_size() for t in host[0])

        def fetch(i):
            with torch.cuda.stream(copy_stream):
                xb = host[i % 2][0].to(device, non_blocking=True)
                yb = host[i % 2][1].to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return xb, yb, ev

        def e2e_loop(n):
            nxt = fetch(0)
            for i in range(n):
                xb, yb, ev = nxt
                torch.cuda.current_stream().wait_event(ev)
                if i + 1 < n:
                    nxt = fetch(i + 1)
                l = one_step(xb, yb)
                xb.record_stream(torch.cuda.current_stream())
                yb.record_stream(torch.cuda.current_stream())
                l.item()                                   # D2H read of the step's result

        e2e_loop(max(1, min(warmup, 2)))
        barrier()
        t0 = time.perf_counter()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        e2e_loop(steps)
        s1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ems = torch.tensor([max(s0.elapsed_time(s1), wall_ms)], device=device)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": world * batch / (ems.item() / steps / 1e3), "unit": w["unit"],
                      "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4}

    # the graph's private pool holds one full set of activations: release it before the eager instrumented steps
    # allocate theirs (UNet 3+ at 512^2 x 8 would not fit three sets in 180 GB)
    if gstep is not None:
        loss = gstep = one_step = None      # (the static loss tensor lives in the graph's pool too)
        release_device_memory()

    # ---- roofline: five instrumented steps (averaged) with per-call CUDA events on the launching stream.
    # Every rank runs it (its backward contains the gradient all-reduce, so a rank-0-only step would wait
    # for peers forever); only rank 0 records and reports.  Per-kernel times must not include the contention
    # of the overlapped weight-gradient stream: the instrumented step runs single-stream.
    prof = KernelProfiler(real_cin=w["cin"], steps=5)
    if rank == 0:
        _lib.profile_hook = prof.hook
    side_was = _engine.WGRAD_STREAM
    _engine.WGRAD_STREAM = False
    with ClockSampler(device.index, period=0.005) as prof_clocks:
        for _ in range(prof.steps):
            eager_step(x, y)
        torch.cuda.synchronize()
    _engine.WGRAD_STREAM = side_was
    _lib.profile_hook = None
    barrier()
    res["roofline"], res["breakdown"] = None, None
    if rank == 0:
        fam = prof.summarize()
        total_ms = sum(f["ms"] for f in fam.values())
        breakdown = {k: {"ms": round(f["ms"], 3), "share": round(f["ms"] / total_ms, 4),
                         "launches": f["launches"],
                         **({"tflops": round(f["flops"] / f["ms"] / 1e9, 1)} if f["flops"] else {})}
                     for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(fam.items(), key=lambda kv: kv[1]["ms"])
        if top[1]["flops"] > 0:
            traffic, tsrc = traffic_for(name, top[0])
            ach = top[1]["flops"] / top[1]["ms"] / 1e9
            # Which measured peak the kernel is held against follows the clock it was timed at: the instrumented
            # steps are eager with an event pair around every call, the GPU idles between kernels and — under the
            # same 1 kW cap — runs them near the maximum SM clock, i.e. in the regime of the BURST cuBLAS figure
            # (the sustained one was taken at a median 1.32 GHz, MEASURED_PEAKS.json).  Below 90 % of the maximum
            # clock the sustained figure applies.
            pc = prof_clocks.summary()
            burst = bool(pc["sm_mhz"] and pc["sm_max_mhz"] and pc["sm_mhz"] >= 0.9 * pc["sm_max_mhz"])
            peak = pk["tf_burst"] if burst else pk["tf_sustained"]
            res["roofline"] = {
                "kernel": top[0], "bound": "tensor", "achieved": round(ach, 1),
                "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
                "frac_of_sustained": round(ach / pk["tf_sustained"], 4), "frac_of_burst": round(ach / pk["tf_burst"], 4),
                "traffic": traffic,
                "traffic_unit": "DRAM bytes per launch",
                "traffic_source": None if tsrc is None else
                f"STATIC: one ncu pass of this workload committed as {tsrc}; not measured in this run",
                "algorithmic_flop_per_launch": round(top[1]["flops"] / top[1]["launches"]),
                "peak_source": pk["source"] + (" (burst cuBLAS bf16: kernels timed at %s of %s MHz)" % (pc["sm_mhz"], pc["sm_max_mhz"])
                                               if burst else " (sustained cuBLAS bf16)"),
                "avg_launch_ms": round(top[1]["ms"] / top[1]["launches"], 4),
                # the instrumented steps are eager, single-stream, with an event pair around every call: the GPU
                # draws less power than in the timed steps and clocks higher under the same cap — compare
                "sm_mhz_instrumented": pc["sm_mhz"], "sm_mhz_timed": clocks.summary()["sm_mhz"],
                "instrumented_step_ms": round(total_ms, 3)}
        breakdown["_step"] = res["step"]
        res["breakdown"] = breakdown
    return res


def release_device_memory():
    """Between workloads: the model / optimizer / tape of the previous one are garbage only after measure()
    has returned — collect them, then hand the cached blocks back so the next workload starts clean."""
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()


EXTRA_WORKLOADS = ("cfg3p", "cfg2")


def run_b200(args, w):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    head = measure(args, args.workload, w, args.steps, args.warmup, device, world, rank, full=True)
    batch = head["batch"]
    release_device_memory()

    # ---- the other two training configurations, device-resident, a few steps each (not the headline)
    extra = None
    if args.extras and args.workload == "cfg3" and not args.batch:
        extra = {}
        for name in EXTRA_WORKLOADS:
            try:
                r = measure(args, name, WORKLOADS[name], max(3, min(args.steps, 6)), 3, device, world, rank,
                            full=False)
                extra[name] = {"workload": WORKLOADS[name]["desc"], "value": round(r["value"], 3),
                               "unit": WORKLOADS[name]["unit"], "ms_per_step": round(r["ms_step"], 3),
                               "per_gpu_batch": r["batch"], "gpu_launches": r["launches"],
                               "step_launch": r["step_launch"], **r["step"], "clocks": r["clocks"]}
            except Exception as e:   # an extra must never take the headline line down with it
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            release_device_memory()

    # ---- reference CPU baseline on a bounded sample, rank 0 at N = 1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_cpu_sample(w, steps=2, warmup=1)

    if rank == 0:
        out = {
            "metric": "train " + w["unit"], "value": round(head["value"], 3), "unit": w["unit"], "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(head["ms_step"], 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": bench_config(args.workload, w, batch, world),
            "e2e": head["e2e"], "gpu_launches": head["launches"], "step_launch": head["step_launch"],
            "clocks": head["clocks"], "roofline": head["roofline"], "cpu_baseline": cpu, "breakdown": head["breakdown"],
            "final_loss": head["final_loss"], "extra": extra,
        }
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def bench_config(name, w, batch, world):
    """`config` of the JSON line — the SAME dict in both arms (the driver compares them)."""
    return {"workload": f"{name}: {w['desc']}", "per_gpu_batch": batch,
            "global_batch": batch * world, "parallelism": f"dp{world}",
            "l2": "per-step inputs + activations far exceed the 126 MB L2 (no flush needed)",
            "streams": "weight gradients on a side stream in the timed steps; the roofline / "
                       "breakdown step runs single-stream and eager",
            "launch": "the whole step is one CUDA-graph replay per rank (GraphedTrainStep), NCCL all-reduces inside",
            "reference_sample": REFERENCE_SAMPLE[w["kind"]]}


# --------------------------------------------------------------------------------------- cfg5 sweep
def run_sweep(args):
    """BASELINE.json configs[4]: inference-only sweep (eval + no_grad; logits out), one GPU.  Per point:
    device-resident latency p50 / p99 over the timed iterations (CUDA events per call), throughput, the
    fraction of the sustained bf16 peak, and the end-to-end latency of one request through the public
    module call with a pinned host input (H2D) and the logits read back (D2H)."""
    import mdeical_image_segmentation_b200 as B
    from mdeical_image_segmentation_b200 import _lib
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pk = peaks()
    torch.manual_seed(0)
    nets = {"2d": B.UNet(1, 1).to(device).eval(), "3d": B.UNet3D(1, 1).to(device).eval()}
    rows = []
    iters = max(args.steps, 5)
    _lib.launch_count = 0
    with ClockSampler(local) as clocks, torch.no_grad():
        for kind, batch, size, tf in SWEEP:
            net = nets[kind]
            g = torch.Generator().manual_seed(batch)
            xh = torch.randn((batch, 1) + size, generator=g).pin_memory()
            x = xh.to(device)
            for _ in range(max(args.warmup, 3)):
                net(x)
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(iters)]
            for e0, e1 in evs:
                e0.record()
                net(x)
                e1.record()
            torch.cuda.synchronize()
            lat = sorted(e0.elapsed_time(e1) for e0, e1 in evs)
            mean = sum(lat) / len(lat)
            # the same forward captured once as a CUDA graph (one launch per request)
            g = B.GraphedInference(net, x)
            for _ in range(3):
                g(x)
            torch.cuda.synchronize()
            gl = []
            for _ in range(iters):
                t0 = time.perf_counter()
                g(x)
                torch.cuda.synchronize()
                gl.append((time.perf_counter() - t0) * 1e3)
            gl.sort()
            del g
            out = net(x)
            oh = torch.empty(out.shape, dtype=out.dtype).pin_memory()
            for it in range(4):          # first trip (allocator warm-up) untimed
                if it == 1:
                    t0 = time.perf_counter()
                out = net(xh.to(device, non_blocking=True))
                oh.copy_(out, non_blocking=True)
                torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
            rows.append({"model": "UNet(1,1)" if kind == "2d" else "UNet3D(1,1) gcr", "batch": batch,
                         "size": "x".join(map(str, size)), "p50_ms": round(lat[len(lat) // 2], 3),
                         "p99_ms": round(lat[min(len(lat) - 1, int(0.99 * len(lat)))], 3),
                         "per_s": round(batch / (mean / 1e3), 2),
                         "graph_p50_ms": round(gl[len(gl) // 2], 3),
                         "graph_p99_ms": round(gl[min(len(gl) - 1, int(0.99 * len(gl)))], 3),
                         "tflops": round(tf * batch / (mean / 1e3), 1),
                         "frac_of_sustained_peak": round(tf * batch / (mean / 1e3) / pk["tf_sustained"], 4),
                         "e2e_ms": round(e2e_ms, 3), "e2e_per_s": round(batch / (e2e_ms / 1e3), 2),
                         "h2d_bytes": xh.numel() * 4, "d2h_bytes": out.numel() * 4})
            del x, xh, out, oh
            torch.cuda.empty_cache()
    serving = serving_rows(device, iters)
    head = next(r for r in rows if r["size"] == "128x128x128" and r["batch"] == 2)
    emit({
        "metric": "inference vols/s", "value": head["per_s"], "unit": "vols/s", "n_gpus": 1, "steps": iters,
        "warmup": max(args.warmup, 3), "ms_per_step": head["p50_ms"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "cfg5: inference-only sweep, UNet(1,1) 256^2-1024^2 and UNet3D(1,1) 96^3-192^3; "
                               "headline value = UNet3D 128^3 batch 2", "l2": "activations exceed L2 except "
                               "at batch 1 256^2 (latency point)"},
        "e2e": {"value": head["e2e_per_s"], "unit": "vols/s", "h2d_bytes_per_step": head["h2d_bytes"],
                "d2h_bytes_per_step": head["d2h_bytes"]},
        "gpu_launches": _lib.launch_count, "clocks": clocks.summary(), "roofline": None, "cpu_baseline": None,
        "sweep": rows, "serving": serving})


def serving_rows(device, iters):
    """Request latency of the serving core behind api/api.py's POST / (serving.SegmentationService): image file on
    disk -> preprocessing (resize 256, centre-crop 224, normalise) -> bucketed CUDA-graph forward of UNet(3, 1) ->
    mask on the device -> PNG written, path returned.  Wall-clock per handle() / handle_many() call, p50 / p99."""
    import tempfile

    import numpy as np
    import mdeical_image_segmentation_b200 as B
    try:
        from PIL import Image
    except ImportError:
        return {"unavailable": "PIL is not installed"}
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        rng = np.random.default_rng(0)
        paths = []
        for i in range(8):
            path = os.path.join(tmp, f"img{i}.png")
            Image.fromarray(rng.integers(0, 255, (300, 400, 3), dtype=np.uint8)).save(path)
            paths.append(path)
        net = B.UNet(3, 1).to(device).eval()
        for graphs in (True, False):
            svc = B.SegmentationService(net, os.path.join(tmp, "masks"), in_channels=3, max_batch=8, use_graphs=graphs)
            for nreq in (1, 8):
                reqs = [{"image_path": p} for p in paths[:nreq]]
                for _ in range(3):
                    ans = svc.handle_many(reqs)
                assert all(a["status"] == 200 and os.path.exists(a["response"]) for a in ans)
                svc.latencies_ms.clear()
                for _ in range(max(iters, 20)):
                    svc.handle_many(reqs)
                lat = svc.latency_percentiles()
                # the device part alone: preprocessed inputs -> masks on the host
                inputs = [svc.build_input(p) for p in paths[:nreq]]
                dev = []
                for _ in range(max(iters, 20)):
                    t0 = time.perf_counter()
                    svc.predict_masks(inputs)
                    dev.append((time.perf_counter() - t0) * 1e3)
                dev.sort()
                rows.append({"requests_per_call": nreq, "cuda_graphs": graphs, "n": lat["n"],
                             "p50_ms": round(lat["p50_ms"], 3), "p99_ms": round(lat["p99_ms"], 3),
                             "requests_per_s": round(nreq / (lat["p50_ms"] / 1e3), 1),
                             "predict_p50_ms": round(dev[len(dev) // 2], 3),
                             "predict_p99_ms": round(dev[min(len(dev) - 1, int(0.99 * len(dev)))], 3)})
    return {"model": "UNet(3,1), 224x224 crop of a 300x400 PNG", "contract": "api/api.py:52-82", "rows": rows}


# --------------------------------------------------------------------------------------- CPU arm
def _cpu_sample_shape(w):
    """Bounded sample of the workload for the host-core arm: (size, scale = sample voxels / workload voxels)."""
    if w["kind"] == "3d":
        size = (64, 64, 64)
    elif w.get("unet_type", "UNet") != "UNet":
        size = (256, 256)
    else:
        size = w["size"]
    vox, full = 1, 1
    for a, b in zip(size, w["size"]):
        vox, full = vox * a, full * b
    return size, vox / full


def _reference_step_fn(w, x, y):
    """fwd + loss of the reference's OWN modules (oracle/_ref, unmodified), or None when they cannot run
    here (copy absent, or the UNet 3+ loss: its pytorch_msssim dependency is not installed)."""
    if w.get("unet_type", "UNet") != "UNet" or w.get("residual") == "se":
        return None
    try:
        from oracle import ref_loader
        if not ref_loader.use_copy():
            return None
        ns = ref_loader.load()
    except Exception as e:   # noqa: BLE001
        print(f"reference copy not importable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        return None
    torch.manual_seed(0)
    if w["kind"] == "2d":
        with contextlib.redirect_stdout(sys.stderr):
            model = ns.U2.UNetModel(ns.U2.UNetConfig(w["cin"], w["cout"], "UNet")).train()
        params = list(model.parameters())
        return (lambda: model(images=x, labels=y).loss), params, torch.optim.AdamW(params, lr=1e-4,
                                                                                    weight_decay=1e-3)
    cls = ns.M3.ResidualUNet3D if w.get("residual") else ns.M3.UNet3D
    kw = {} if w.get("residual") else {"upsample": w["upsample"]}
    net = cls(w["cin"], w["cout"], layer_order=w["order"], **kw).train()
    crit = ns.L3.DiceLoss(normalization="sigmoid") if w.get("loss") == "dice" else ns.L3.BCEDiceLoss(1.0, 1.0)
    params = list(net.parameters())
    return (lambda: crit(net(x), y)), params, torch.optim.Adam(params, lr=1e-4)


def _port_step_fn(w, x, y):
    """The same step through the oracle port (oracle/unet.py, pinned to the reference by tests/golden)."""
    from oracle import unet as O
    import mdeical_image_segmentation_b200 as B
    torch.manual_seed(0)

    def leaves(net):
        return {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
                for k, v in net.state_dict().items()}

    if w.get("unet_type", "UNet") != "UNet":       # UNet 3+ family
        from oracle import unet3plus as O3
        from oracle.msssim import segmentation_loss
        deep = w["unet_type"] == "UNet_3Plus_DeepSup"
        sd = leaves(getattr(B, w["unet_type"])(w["cin"], w["cout"]))

        def fwd():
            out = O3.unet3plus_forward(sd, x, training=True, deep_sup=deep)
            return sum(segmentation_loss(d, y) for d in out) if deep else segmentation_loss(out, y)
        params = [v for v in sd.values() if v.requires_grad]
        return fwd, params, torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-3)
    if w["kind"] == "2d":
        sd = leaves(B.UNet(w["cin"], w["cout"]))
        params = list(sd.values())
        return (lambda: O.unet_model_loss(O.unet2d_forward(sd, x), y, w["cout"])), params, \
            torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-3)
    if w.get("residual"):
        net = (B.ResidualUNetSE3D if w["residual"] == "se" else B.ResidualUNet3D)(w["cin"], w["cout"],
                                                                                  layer_order=w["order"])
    else:
        net = B.UNet3D(w["cin"], w["cout"], layer_order=w["order"], upsample=w["upsample"])
    sd = leaves(net)
    lossf = O.dice_loss if w.get("loss") == "dice" else O.bce_dice_loss
    if w.get("residual"):
        fwd = lambda: lossf(O.residual_unet3d_forward(sd, x, layer_order=w["order"]), y)  # noqa: E731
    else:
        fwd = lambda: lossf(O.unet3d_forward(sd, x, layer_order=w["order"], upsample=w["upsample"],  # noqa: E731
                                             training=True), y)
    params = [v for v in sd.values() if v.requires_grad]
    return fwd, params, torch.optim.Adam(params, lr=1e-4)


def run_cpu_sample(w, steps, warmup, batch=None):
    """The reference path on the host cores: fwd + loss + bwd + optimizer step on a bounded sample (3D: the
    per-GPU batch of 64^3 crops per step; 2D: one image per step; throughput scaled by voxel count)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    size, frac = _cpu_sample_shape(w)
    b = batch or (w["batch"] if w["kind"] == "3d" else 1)
    x, y = make_batch(dict(w, size=size), b, 7)
    made = _reference_step_fn(w, x, y)
    kind = "reference"
    if made is None:
        made, kind = _port_step_fn(w, x, y), "port"
    fwd, _, opt = made
    dims = "x".join(str(v) for v in size)
    sample = f"{b} sample(s) of {dims} per step"
    if frac != 1.0:
        sample += f", scaled by voxel count to {'x'.join(str(v) for v in w['size'])}"

    def step():
        opt.zero_grad()
        l = fwd()
        l.backward()
        opt.step()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return {"value": round(b * frac / dt, 5), "unit": w["unit"], "cores": cores, "kind": kind,
            "sample": sample, "s_per_step": round(dt, 3)}


def run_reference(args, w):
    """`--impl reference`: --steps / --warmup are honoured as given (a step = the bounded sample above)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = run_cpu_sample(w, steps=max(1, args.steps), warmup=max(0, args.warmup))
    out = {"impl": "reference", "metric": "train " + w["unit"], "value": cpu["value"], "unit": w["unit"],
           "n_gpus": world, "steps": max(1, args.steps), "warmup": max(0, args.warmup),
           "ms_per_step": cpu["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": bench_config(args.workload, w, args.batch or w["batch"], world),
           "cpu_baseline": cpu, "gpu_launches": 0,
           "e2e": {"value": cpu["value"], "unit": w["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS) + ["cfg5"])
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the short cfg2 / cfg3p measurements nested under 'extra'")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="time eager steps (default: one CUDA-graph replay per step, at every N)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.workload == "cfg5":
        _claim_stdout()
        if args.impl == "reference":
            emit({"impl": "reference", "unavailable": "cfg5 is a sweep; use cfg2/cfg3 for the CPU arm"})
            return
        return run_sweep(args)
    w = WORKLOADS[args.workload]
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl == "b200":
        # convenience: re-launch under torch.distributed.run, one rank per GPU
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
                                   "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:])
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
