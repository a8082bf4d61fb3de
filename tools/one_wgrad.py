"""Run one conv weight-gradient a few times (for ncu captures): python tools/one_wgrad.py n d h w cin cout"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import ops
from mdeical_image_segmentation_b200.ops import CL
n, d, h, w, cin, cout = map(int, sys.argv[1:7])
x = CL(torch.randn((n, d, h, w, cin), device="cuda").to(torch.bfloat16))
dy = CL(torch.randn((n, d, h, w, cout), device="cuda").to(torch.bfloat16))
k = (3, 3, 3) if d > 1 else (1, 3, 3)
shape = (cout, cin, 3, 3, 3) if d > 1 else (cout, cin, 3, 3)
for _ in range(4):
    ops.conv_bwd_weight(x, dy, k, shape, impl=1)
torch.cuda.synchronize()
