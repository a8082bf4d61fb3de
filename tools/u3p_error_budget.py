"""Where the UNet 3+ parameter gradient leaves the fp32 reference: per-site bf16-storage error budget (CPU only).

The fp32 oracle trains the full 27 M-parameter UNet_3Plus(3, 1) for 30 Adam steps on the blob task, then the gradient
of BCE is taken (a) in fp32 and (b) with ONE class of tensors rounded to bf16 where the B200 path stores bf16 (straight-
through gradient), on a 96 x 128 and on a 256 x 256 batch.  Output: relative L2 distance of the concatenated parameter
gradient to the fp32 one, per site.  python tools/u3p_error_budget.py > profiles/r02_u3p_error_budget.txt"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mdeical_image_segmentation_b200 as B  # noqa: E402
from oracle import unet3plus as O3  # noqa: E402
from test_gpu_parity import _blob_task  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(os.cpu_count() or 1)
net = B.UNet_3Plus(3, 1)
sd = {k: v.clone() for k, v in net.state_dict().items()}
names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
for k in names:
    sd[k].requires_grad_(True)


def task(seed, shape):
    x, m = _blob_task((shape[0], 1) + shape[2:], seed)
    g = torch.Generator().manual_seed(seed + 5000)
    return torch.cat([x, 0.5 * torch.randn((shape[0], 2) + shape[2:], generator=g)], 1), m


bce = F.binary_cross_entropy_with_logits
opt = torch.optim.Adam([sd[k] for k in names], lr=2e-4)
for s in range(30):
    x, m = task(100 + s, (2, 3, 32, 48))
    opt.zero_grad()
    bce(O3.unet3plus_forward(sd, x, True), m).backward()
    opt.step()
trained = {k: v.detach().clone() for k, v in sd.items()}


def group_of(k):
    """conv1..conv5 = encoder levels; decoder stage s = its five branch convs + the fusion conv (unet.py:387-501)."""
    head = k.split(".")[0]
    if "_hd" in head:
        return "decoder stage " + head.split("_hd")[1][0]
    if head.endswith("d_1"):
        return "decoder stage " + head[-4]
    return head


def grad(x, m, sites):
    s = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in trained.items()}
    lo = O3.unet3plus_forward(s, x, True, emulate_bf16=sites)
    bce(lo, m).backward()
    return lo.detach(), torch.cat([s[k].grad.flatten() for k in names])


print("# UNet_3Plus(3,1), 30 Adam steps by the fp32 oracle; rel-L2 of the concatenated parameter gradient to fp32")
print("# when ONE class of stored tensors is rounded to bf16 (fp32 accumulation everywhere)")
for shape in ((2, 3, 96, 128), (2, 3, 256, 256)):
    x, m = task(999, shape)
    lo0, g0 = grad(x, m, False)
    print(f"\nbatch {shape}: |grad| = {g0.norm():.4e}")
    for label, sites in (("filter packs (w)", {"w"}), ("conv inputs (x)", {"x"}), ("pre-BatchNorm conv outputs (pre_bn)", {"pre_bn"}),
                         ("normalised + ReLU outputs (act)", {"act"}), ("bilinear outputs (up)", {"up"}),
                         ("output convs (head)", {"head"}), ("all but pre_bn", O3.ALL_SITES - {"pre_bn"}),
                         ("ALL sites (= the product's storage policy)", True)):
        lo, g = grad(x, m, sites)
        print(f"  {label:46s} gradient rel-L2 {((g - g0).norm() / g0.norm()).item():.4f}   "
              f"cosine {F.cosine_similarity(g.double(), g0.double(), 0).item():.5f}   "
              f"logits rel-L2 {((lo - lo0).norm() / lo0.norm()).item():.2e}")
    # the reference's own mixed-precision recipe for comparison: torch.autocast(bf16) around the fp32 oracle forward
    s_ac = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in trained.items()}
    with torch.autocast("cpu", dtype=torch.bfloat16):
        lo = O3.unet3plus_forward(s_ac, x, True)
    bce(lo.float(), m).backward()
    g_ac = torch.cat([s_ac[k].grad.flatten() for k in names])
    print(f"  {'torch.autocast(bf16) of the fp32 network':46s} gradient rel-L2 {((g_ac - g0).norm() / g0.norm()).item():.4f}   "
          f"cosine {F.cosine_similarity(g_ac.double(), g0.double(), 0).item():.5f}   "
          f"logits rel-L2 {((lo.float() - lo0).norm() / lo0.norm()).item():.2e}")
    # where in the network the ALL-sites error sits: share of |g - g0|^2 and own rel-L2 per parameter group
    sizes = [trained[k].numel() for k in names]
    off, groups = 0, {}
    for k, n in zip(names, sizes):
        key = group_of(k)
        d, r = groups.setdefault(key, [0.0, 0.0])
        groups[key] = [d + (g[off:off + n] - g0[off:off + n]).pow(2).sum().item(), r + g0[off:off + n].pow(2).sum().item()]
        off += n
    tot = sum(v[0] for v in groups.values())
    print("  ALL sites, by parameter group:  group | share of squared error | own rel-L2 | share of |grad|^2")
    tg = sum(v[1] for v in groups.values())
    for key, (d, r) in sorted(groups.items()):
        print(f"    {key:22s} {d / tot:6.3f}   {(d / max(r, 1e-30)) ** 0.5:7.4f}   {r / tg:6.3f}")
