"""A/B the row-block count R of the halo kernel on chosen layer shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import ops, _lib
from mdeical_image_segmentation_b200.ops import CL
dev = "cuda"
shapes = [("2D 64->64 512^2 b32", 32, 1, 512, 512, 64, 64, False), ("2D 128->64 512^2 b32", 32, 1, 512, 512, 128, 64, False),
          ("2D 64->128 256^2 b32", 32, 1, 256, 256, 64, 128, False), ("2D 128->128 256^2 b32", 32, 1, 256, 256, 128, 128, False),
          ("3D 64->64 128^3 b2", 2, 128, 128, 128, 64, 64, True), ("3D 32->64 128^3 b2", 2, 128, 128, 128, 32, 64, True),
          ("3D 64->128 64^3 b2", 2, 64, 64, 64, 64, 128, True), ("3D 128->128 64^3", 2, 64, 64, 64, 128, 128, True)]
for (name, n, d, h, w, cin, cout, is3d) in shapes:
    x = CL(torch.randn((n, d, h, w, cin), device=dev).to(torch.bfloat16))
    wt = torch.randn((cout, cin, 3, 3, 3) if is3d else (cout, cin, 3, 3), device=dev) * 0.02
    fwd, bwd, k = ops.conv_repack_weight(wt)
    out = CL.empty(n, d, h, w, cout, dev)
    flops = 2.0 * n * d * h * w * cin * cout * (27 if is3d else 9)
    res = []
    for R in (1, 2, 4):
        _lib.lib().b200_set_tuning(2, R)
        for _ in range(2):
            ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res.append(f"R<={R}: {ms:.3f} ms {flops/ms/1e9:.0f} TF/s")
    _lib.lib().b200_set_tuning(2, 4)
    print(name, " | ".join(res), flush=True)
