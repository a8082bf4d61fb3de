"""Timing probe for the im2col first layer: the pieces at cfg3 / cfg2 sizes, epilogue knobs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import ops, _lib  # noqa: E402
from mdeical_image_segmentation_b200.ops import CL  # noqa: E402

dev = torch.device("cuda", 0)


def t(fn, it=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


for name, n, d, h, w, cin, cout, kd in [("cfg3 1->32 128^3 b2", 2, 128, 128, 128, 1, 32, 3),
                                        ("cfg2 3->64 512^2 b64", 64, 1, 512, 512, 3, 64, 1)]:
    xb = torch.zeros((n, d, h, w, 16), dtype=torch.bfloat16, device=dev)
    xb[..., :cin] = torch.randn((n, d, h, w, cin), device=dev).to(torch.bfloat16)
    x = CL(xb, cin, 0, 16 - cin)
    kp = (9 * kd * cin + 15) // 16 * 16
    a = ops.im2col(x, kd, kp)
    wt = torch.randn((cout, kp, 1, 1, 1), device=dev) * 0.1
    fwd, bwd, k1 = ops.conv_repack_weight(wt)
    out = CL.empty(n, d, h, w, cout, dev)
    dy = CL(torch.randn((n, d, h, w, cout), device=dev).to(torch.bfloat16))
    print(f"== {name} kp={kp}")
    print(f"im2col        {t(lambda: ops.im2col(x, kd, kp)):.3f} ms")
    for ew in (0, 4, 8):
        _lib.lib().b200_set_tuning(5, ew)
        print(f"epi_warps={ew}: fwd {t(lambda: ops.conv_fwd(a, fwd, k1, None, True, out=out, impl=1)):.3f} ms | "
              f"fwd+stats {t(lambda: ops.conv_fwd(a, fwd, k1, None, True, out=out, impl=1, stats=True)):.3f} ms | "
              f"dgrad {t(lambda: ops.conv_bwd_data(dy, bwd, k1, impl=1)):.3f} ms")
    _lib.lib().b200_set_tuning(5, 0)
    for dbg in (1, 2):
        _lib.lib().b200_set_tuning(4, dbg)
        print(f"dbg={dbg}: fwd {t(lambda: ops.conv_fwd(a, fwd, k1, None, True, out=out, impl=1)):.3f} ms")
    _lib.lib().b200_set_tuning(4, 0)
    print(f"wgrad         {t(lambda: ops.conv_bwd_weight(a, dy, k1, wt.shape, impl=1)):.3f} ms")
    # old path for comparison: padded 16-channel 3x3(x3) conv
    ks = (3, 3, 3) if kd == 3 else (3, 3)
    w16 = torch.randn((cout, 16) + ks, device=dev) * 0.1
    f16, b16, k3 = ops.conv_repack_weight(w16)
    xw = x.widened()
    print(f"old fwd {t(lambda: ops.conv_fwd(xw, f16, k3, None, True, out=out, impl=1)):.3f} ms | "
          f"old fwd+stats {t(lambda: ops.conv_fwd(xw, f16, k3, None, True, out=out, impl=1, stats=True)):.3f} ms | "
          f"old dgrad {t(lambda: ops.conv_bwd_data(dy, b16, k3, impl=1)):.3f} ms | "
          f"old wgrad {t(lambda: ops.conv_bwd_weight(xw, dy, k3, w16.shape, impl=1)):.3f} ms")
