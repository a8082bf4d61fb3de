"""Run one conv layer a few times (for ncu captures): python tools/one_layer.py n d h w cin cout [fwd|dgrad] [stats]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import ops
from mdeical_image_segmentation_b200.ops import CL
n, d, h, w, cin, cout = map(int, sys.argv[1:7])
mode = sys.argv[7] if len(sys.argv) > 7 else "fwd"
stats = len(sys.argv) > 8
x = CL(torch.randn((n, d, h, w, cin), device="cuda").to(torch.bfloat16))
wt = torch.randn((cout, cin, 3, 3, 3) if d > 1 else (cout, cin, 3, 3), device="cuda") * 0.02
fwd, bwd, k = ops.conv_repack_weight(wt)
out = CL.empty(n, d, h, w, cout, "cuda")
dy = CL(torch.randn((n, d, h, w, cout), device="cuda").to(torch.bfloat16))
for _ in range(4):
    if mode == "fwd":
        ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1, stats=stats)
    else:
        ops.conv_bwd_data(dy, bwd, k, mask=x, impl=1, stats=stats)
torch.cuda.synchronize()
