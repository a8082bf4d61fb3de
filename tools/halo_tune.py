"""Times forward / data-gradient launches of single layers under tuning-knob settings (CUDA events, steady
state): python tools/halo_tune.py [layer-set].  Not part of the product path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import _lib, ops  # noqa: E402
from mdeical_image_segmentation_b200.ops import CL  # noqa: E402

LAYERS = {   # name: n, d, h, w, cin, cout
    "192_64": (2, 128, 128, 128, 192, 64),
    "64_64": (2, 128, 128, 128, 64, 64),
    "384_128": (2, 64, 64, 64, 384, 128),
    "128_128": (2, 64, 64, 64, 128, 128),
    "768_256": (2, 32, 32, 32, 768, 256),
    "256_256": (2, 32, 32, 32, 256, 256),
    "32_64": (2, 128, 128, 128, 32, 64),
    "16_32": (2, 128, 128, 128, 16, 32),
    "64_64_64": (2, 64, 64, 64, 64, 64),
    "64_128": (2, 64, 64, 64, 64, 128),
    "256_512": (2, 16, 16, 16, 256, 512),
    "2d_128_64": (16, 1, 512, 512, 128, 64),
    "2d_256_128": (16, 1, 256, 256, 256, 128),
    "2d_512_256": (16, 1, 128, 128, 512, 256),
    "2d_1024_512": (16, 1, 64, 64, 1024, 512),
}
DEFAULTS = {0: 1, 1: 1, 2: 4, 3: 1, 4: 0, 5: 0, 6: 1, 7: 0, 8: 1, 9: 80, 10: 0, 11: 0}
lib = _lib.lib()


def set_knobs(pairs):
    for k, v in DEFAULTS.items():
        lib.b200_set_tuning(k, v)
    for k, v in pairs:
        lib.b200_set_tuning(k, v)


def timeit(fn, iters=12):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else list(LAYERS)
    configs = [("default", ())]
    for arg in sys.argv[2:]:          # e.g. "7:128" or "7:128,2:2"
        configs.append((arg, tuple(tuple(int(v) for v in kv.split(":")) for kv in arg.split(","))))
    for name in names:
        n, d, h, w, cin, cout = LAYERS[name]
        x = CL(torch.randn((n, d, h, w, cin), device="cuda").to(torch.bfloat16))
        dy = CL(torch.randn((n, d, h, w, cout), device="cuda").to(torch.bfloat16))
        wt = torch.randn((cout, cin, 3, 3, 3) if d > 1 else (cout, cin, 3, 3), device="cuda") * 0.02
        fwd, bwd, k = ops.conv_repack_weight(wt)
        out = CL.empty(n, d, h, w, cout, "cuda")
        dx = CL.empty(n, d, h, w, cin, "cuda")
        fl = 2.0 * n * d * h * w * cin * cout * (27 if d > 1 else 9)
        which = os.environ.get("HALO_TUNE_OPS", "fwd,dgrad,wgrad").split(",")
        fns = {"fwd": lambda: ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1),
               "dgrad": lambda: ops.conv_bwd_data(dy, bwd, k, mask=x, out=dx, impl=1),
               "wgrad": lambda: ops.conv_bwd_weight(x, dy, k, tuple(wt.shape), impl=1)}
        res = {(c, o): [] for c, _ in configs for o in which}
        for rnd in range(7):                      # interleave the configurations: the power state drifts
            for cname, pairs in configs:
                set_knobs(pairs)
                for o in which:
                    res[(cname, o)].append(timeit(fns[o], iters=5))
        for cname, _ in configs:
            line = f"{name:12s} {cname:16s}"
            for o in which:
                t = sorted(res[(cname, o)])[len(res[(cname, o)]) // 2]
                line += f" | {o} {t:7.3f} ms {fl / t / 1e9:7.1f} TF/s"
            print(line, flush=True)
        set_knobs(())
        del x, dy, out, dx


if __name__ == "__main__":
    main()
