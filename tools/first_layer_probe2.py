"""Times (CUDA events) the CUDA-core first-layer kernels against the tensor-core path at 128^3 x 2; under ncu it
just launches each kernel a few times.  python tools/first_layer_probe2.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import ops  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
g = torch.Generator().manual_seed(0)
x = ops.from_nchw(torch.randn((2, 1, 128, 128, 128), generator=g).cuda(), pad=True)
w = (0.2 * torch.randn((32, 1, 3, 3, 3), generator=g)).cuda()
b = torch.randn(32, generator=g).cuda()
dy = ops.from_nchw(torch.randn((2, 32, 128, 128, 128), generator=g).cuda())


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("conv3d_c1_fwd (stats)      %.3f ms" % timed(lambda: ops.conv3d_c1_fwd(x, w, b, False, stats=True)))
print("conv3d_c1_fwd (no stats)   %.3f ms" % timed(lambda: ops.conv3d_c1_fwd(x, w, b, False, stats=False)))
wp = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, 0, 0, 15))
wp_fwd, wp_bwd, k = ops.conv_repack_weight(wp, need_bwd=True)
xw = x.widened()
print("tensor-core fwd (16 ch)    %.3f ms" % timed(lambda: ops.conv_fwd(xw, wp_fwd, k, b, False, impl=1, stats=True)))
print("im2col K'=64 + GEMM        %.3f ms" % timed(lambda: ops.conv_bwd_weight(
    ops.im2col(ops.CL(x.buf, 1, 0, 15), 3, 64, ones=True), dy, (1, 1, 1), (32, 64, 1, 1, 1), impl=1)))
