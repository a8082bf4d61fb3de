"""Per-call timing of one training step (CUDA events around every C-ABI call): which layers are slow.
Usage: python tools/layer_profile.py cfg3|cfg2|cfg3p"""
import contextlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mdeical_image_segmentation_b200 import _lib  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
for kv in os.environ.get("B200_TUNING", "").split(","):   # e.g. B200_TUNING=6:0 (tuning key:value pairs)
    if kv:
        k, v = kv.split(":")
        _lib.lib().b200_set_tuning(int(k), int(v))
w = bench.WORKLOADS[wl]
dev = torch.device("cuda", 0)
model, opt, step_fn = bench.build_model(w, dev)
x, y = bench.make_batch(w, w["batch"], 1, dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = step_fn(x, y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
recs = []


@contextlib.contextmanager
def hook(name, args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    yield
    e1.record()
    d = args[0]._obj if hasattr(args[0], "_obj") else None
    desc = ""
    if d is not None and hasattr(d, "c_in"):
        dims = f"{d.n}x{getattr(d, 'd', 1)}x{d.h}x{d.w}"
        desc = f"{d.c_in:>4}->{d.c_out:<4} {dims}"
    impl = _lib.lib().b200_last_impl() if "conv" in name else 0
    recs.append((name.replace("b200_", ""), desc, bench.conv_flops(name, args), impl, e0, e1))


_lib.profile_hook = hook
step()
_lib.profile_hook = None
torch.cuda.synchronize()
tot = 0.0
for name, desc, fl, impl, e0, e1 in recs:
    ms = e0.elapsed_time(e1)
    tot += ms
    if ms < 0.15 and "repack" in name:
        continue
    tf = f"{fl / ms / 1e9:7.0f} TF/s" if fl else ""
    print(f"{name:28s} {desc:34s} impl={impl} {ms:8.3f} ms {tf}")
print(f"total {tot:.2f} ms")
