"""Attribute the time of the forward / dgrad GEMM kernels on the epilogue-bound 2D layers:
knobs = weight-resident on/off, row-blocks R, epilogue debug mode (0 full, 1 no global stores,
2 no epilogue work at all)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import ops, _lib
from mdeical_image_segmentation_b200.ops import CL
dev = "cuda"
L = _lib.lib()
shapes = [("3D 64->64 128^3 b2", 2, 128, 128, 128, 64, 64), ("3D 192->64 128^3 b2", 2, 128, 128, 128, 192, 64),
          ("3D 32->64 128^3 b2", 2, 128, 128, 128, 32, 64), ("3D 128->128 64^3 b2", 2, 64, 64, 64, 128, 128),
          ("2D 64->64 512^2 b32", 32, 1, 512, 512, 64, 64), ("2D 128->64 512^2 b32", 32, 1, 512, 512, 128, 64),
          ("2D 64->128 256^2 b32", 32, 1, 256, 256, 64, 128), ("2D 128->128 256^2 b32", 32, 1, 256, 256, 128, 128)]


def timeit(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5


for (name, n, d, h, w, cin, cout) in shapes:
    x = CL(torch.randn((n, d, h, w, cin), device=dev).to(torch.bfloat16))
    wt = torch.randn((cout, cin, 3, 3, 3) if d > 1 else (cout, cin, 3, 3), device=dev) * 0.02
    fwd, bwd, k = ops.conv_repack_weight(wt)
    out = CL.empty(n, d, h, w, cout, dev)
    mask = CL(torch.randn((n, d, h, w, cout), device=dev).to(torch.bfloat16))
    flops = 2.0 * n * d * h * w * cin * cout * (27 if d > 1 else 9)
    dy = CL(torch.randn((n, d, h, w, cout), device=dev).to(torch.bfloat16))
    for resident in (1,):
        for ew in (4, 8):
            L.b200_set_tuning(3, resident); L.b200_set_tuning(5, ew)
            row = []
            for dbg in (0, 2):
                L.b200_set_tuning(4, dbg)
                ms = timeit(lambda: ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1))
                row.append(f"dbg{dbg} {ms:.3f}ms {flops/ms/1e9:5.0f}TF")
            L.b200_set_tuning(4, 0)
            ms = timeit(lambda: ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1, stats=True))
            row.append(f"fwd+sq {ms:.3f}ms")
            ms = timeit(lambda: ops.conv_bwd_data(dy, bwd, k, mask=x, impl=1))
            row.append(f"dgrad+mask {ms:.3f}ms")
            ms = timeit(lambda: ops.conv_bwd_data(dy, bwd, k, mask=x, impl=1, stats="sum"))
            row.append(f"dgrad+mask+sum {ms:.3f}ms")
            ms = timeit(lambda: ops.conv_bwd_data(dy, bwd, k, impl=1, stats=True, dot=x))
            row.append(f"dgrad+dot {ms:.3f}ms")
            print(f"{name} resident={resident} epi_warps={ew}: " + " | ".join(row), flush=True)
    L.b200_set_tuning(3, 1); L.b200_set_tuning(5, 0); L.b200_set_tuning(4, 0)
# ConvTranspose 2x2 forward (scatter epilogue) and backward-data
for (name, n, h, w, cin, cout) in [("convT 128->64 256^2 b32", 32, 256, 256, 128, 64), ("convT 256->128 128^2 b32", 32, 128, 128, 256, 128)]:
    x = CL(torch.randn((n, 1, h, w, cin), device=dev).to(torch.bfloat16))
    wt = torch.randn((cin, cout, 2, 2), device=dev) * 0.02
    wf, wb = ops.convt_repack_weight(wt)
    out = CL.empty(n, 1, 2 * h, 2 * w, cout, dev)
    dy = CL(torch.randn((n, 1, 2 * h, 2 * w, cout), device=dev).to(torch.bfloat16))
    flops = 2.0 * n * h * w * cin * cout * 4
    for resident in (1,):
      for ew in (4, 8):
        L.b200_set_tuning(3, resident); L.b200_set_tuning(5, ew)
        row = []
        for dbg in (0, 2):
            L.b200_set_tuning(4, dbg)
            ms = timeit(lambda: ops.convt_fwd(x, wf, None, out=out, impl=1))
            ms2 = timeit(lambda: ops.convt_bwd_data(dy, wb, mask=x, impl=1, stats="sum"))
            row.append(f"dbg{dbg} fwd {ms:.3f}ms {flops/ms/1e9:5.0f}TF dgrad+mask+sum {ms2:.3f}ms")
        L.b200_set_tuning(4, 0)
        print(f"{name} resident={resident} epi_warps={ew}: " + " | ".join(row), flush=True)
L.b200_set_tuning(3, 1); L.b200_set_tuning(5, 0)
