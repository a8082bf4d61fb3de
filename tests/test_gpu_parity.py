"""GPU parity tests (run on the B200 box): the CUDA path, called through the C ABI, against the CPU
oracle and the reference's golden vectors.

Tolerances (BASELINE.json north_star): logits and gradients within 2e-2 relative (bf16 storage, fp32
accumulate), masks agreeing on >= 99.9 % of voxels, Dice/IoU within 1e-3 — asserted on briefly
trained weights with signal-bearing inputs, where an ideal bf16 pipeline can meet them;
random-init fixtures are held to the looser bound stated at each assert.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2, sd_from
from oracle import metrics as OM
from oracle import unet as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    import mdeical_image_segmentation_b200 as B
    return B


def _grads(net):
    return {k: p.grad.detach().cpu() for k, p in net.named_parameters() if p.grad is not None}


def _global_rel(ga, gb, keys):
    a = torch.cat([ga[k].flatten().double() for k in keys])
    b = torch.cat([gb[k].flatten().double() for k in keys])
    return ((a - b).norm() / b.norm()).item()


# ----------------------------------------------------------------------------------- golden: 2D
@pytest.mark.parametrize("name,cin,cout", [("unet2d_1_1", 1, 1), ("unet2d_3_4", 3, 4)])
def test_unet2d_vs_reference_golden(B, name, cin, cout):
    fix = load_golden(name)
    torch.manual_seed(int(fix["seed"]))
    model = B.UNetModel(B.UNetConfig(cin, cout)).cuda()
    # seeded init is consumed by UNet's convs first (same order as the reference)
    x = torch.from_numpy(fix["x"]).cuda()
    y = torch.from_numpy(fix["y"]).cuda()
    out = model(images=x, labels=y)
    # random init + noise input, fp32 reference: bf16 pipeline measured ~2e-3; bound 1e-2
    assert rel_l2(out.logits.cpu(), fix["logits"]) < 1e-2
    assert abs(out.loss.item() - float(fix["loss"])) < 1e-3
    out.loss.backward()
    g = {k[len("unet."):]: v for k, v in _grads(model).items()}
    names = [str(s) for s in fix["grad_names"]]
    l2 = np.array([g[k].norm().item() for k in names])
    ref = fix["grad_l2"]
    # per-tensor gradient norms: within 5 % or an absolute floor tied to the global norm
    floor = 2e-3 * np.linalg.norm(ref)
    assert np.all(np.abs(l2 - ref) < 0.05 * ref + floor), np.abs(l2 - ref) / (ref + floor)
    # random init + noise is the regime where bf16 storage costs the most (19-29 % on the
    # 3D nets); the strict 2e-2 gradient bound is asserted on trained weights further down
    for k, v in fix.items():
        if k.startswith("grad__"):
            assert rel_l2(g[k[6:]], v) < 0.2, (k, rel_l2(g[k[6:]], v))


# ----------------------------------------------------------------------------------- golden: 3D family
@pytest.mark.parametrize("name", ["unet3d_gcr_nearest", "unet3d_cbr_deconv", "unet3d_cr_nearest",
                                  "unet2d3_gcr_nearest", "unet2d3_cbr_deconv"])
def test_unet3d_family_vs_reference_golden(B, name):
    fix = load_golden(name)
    is3d = bool(fix["is3d"])
    cls = B.UNet3D if is3d else B.UNet2D
    net = cls(2, 2, f_maps=[8, 16, 32], num_groups=4, layer_order=str(fix["layer_order"]),
              upsample=str(fix["upsample"]))
    net.load_state_dict(sd_from(fix), strict=True)
    net = net.cuda().train()
    x, t = torch.from_numpy(fix["x"]).cuda(), torch.from_numpy(fix["t"]).cuda()
    logits = net(x)
    # random init + noise: an ideal bf16 pipeline sits at 4-7e-2 here; bound 1e-1
    assert rel_l2(logits.cpu(), fix["logits"]) < 1e-1
    loss = B.BCEDiceLoss(1.0, 1.0)(logits, t)
    assert abs(loss.item() - float(fix["loss"])) < 2e-2
    loss.backward()
    g = _grads(net)
    ref = {k[6:]: torch.from_numpy(v) for k, v in fix.items() if k.startswith("grad__")}
    assert set(g) == set(ref)
    assert _global_rel(g, ref, sorted(ref)) < 0.35          # random-init gradients: 19-29 %
    for k, v in fix.items():
        if k.startswith("after__"):                          # BatchNorm running statistics
            assert rel_l2(net.state_dict()[k[7:]].cpu(), v) < 2e-2, k
    net.eval()
    with torch.no_grad():
        ev = net(x)
    assert rel_l2(ev.cpu(), fix["logits_eval"]) < 1e-1


def test_unetconv2_block_vs_reference_golden(B):
    fix = load_golden("unetconv2_bn")
    blk = B.unetConv2(3, 8, True)
    blk.load_state_dict(sd_from(fix))
    blk = blk.cuda().train()
    x = torch.from_numpy(fix["x"]).cuda().requires_grad_(True)
    y = blk(x)
    assert rel_l2(y.cpu(), fix["y"]) < 1e-2
    y.square().sum().backward()
    assert rel_l2(x.grad.cpu(), fix["dx"]) < 3e-2
    named = dict(blk.named_parameters())
    gnorm = np.sqrt(sum(float(np.sum(v.astype(np.float64) ** 2)) for k, v in fix.items()
                        if k.startswith("grad__")))
    for k, v in fix.items():
        if k.startswith("grad__"):
            # conv biases feeding a BatchNorm have an analytically zero gradient: absolute floor
            err = (named[k[6:]].grad.cpu() - torch.from_numpy(v)).norm().item()
            assert err < 3e-2 * np.linalg.norm(v) + 2e-3 * gnorm, (k, err, np.linalg.norm(v), gnorm)


# ----------------------------------------------------------------------------------- trained fixtures
def _blob_task(shape, seed):
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(shape, generator=g)
    pool = F.avg_pool3d if len(shape) == 5 else F.avg_pool2d
    blur = pool(noise, 5, stride=1, padding=2)
    mask = (blur > torch.quantile(blur.flatten(), 0.7)).float()
    x = 0.5 * torch.randn(shape, generator=g) + mask
    return x, mask


def _train_oracle(sd, fwd, loss_fn, shape, steps, lr=1e-3):
    params = [v for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    for p in params:
        p.requires_grad_(True)
    opt = torch.optim.Adam(params, lr=lr)
    for s in range(steps):
        x, m = _blob_task(shape, 100 + s)
        opt.zero_grad()
        loss_fn(fwd(sd, x), m).backward()
        opt.step()
    return sd


@pytest.mark.parametrize("order,upsample", [("gcr", "default"), ("cbr", "deconv")])
def test_unet3d_trained_parity(B, order, upsample):
    torch.manual_seed(0)
    torch.set_num_threads(max(1, min(8, torch.get_num_threads())))
    net = B.UNet3D(1, 1, f_maps=[16, 32, 64], layer_order=order, upsample=upsample)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    kw = dict(num_levels=3, layer_order=order, num_groups=8, upsample=upsample, is3d=True)
    shape = (2, 1, 32, 32, 32)
    fwd = lambda s, x: O.unet3d_forward(s, x, training=True, **kw)  # noqa: E731
    sd = _train_oracle(sd, fwd, O.bce_dice_loss, shape, steps=40)
    x, m = _blob_task(shape, 999)
    for v in sd.values():
        if v.grad is not None:
            v.grad = None
    sd_eval = {k: v.detach().clone() for k, v in sd.items()}
    ref_logits = fwd(sd, x)
    ref_loss = O.bce_dice_loss(ref_logits, m)
    ref_loss.backward()
    ref_g = {k: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}

    net.load_state_dict(sd_eval)
    net = net.cuda().train()
    logits = net(x.cuda())
    loss = B.BCEDiceLoss(1.0, 1.0)(logits, m.cuda())
    loss.backward()
    lg = logits.detach().cpu()
    assert rel_l2(lg, ref_logits.detach()) < 2e-2
    assert (lg - ref_logits.detach()).abs().max() / ref_logits.detach().abs().max() < 2e-2
    agree = ((lg > 0) == (ref_logits.detach() > 0)).float().mean().item()
    assert agree >= 0.999, agree
    assert abs(loss.item() - ref_loss.item()) < 1e-3
    # Dice / IoU metrics within 1e-3
    d_ref = O.dice_coefficient(torch.sigmoid(ref_logits.detach()), m).item()
    d_dev = B.metrics.DiceCoefficient()(torch.sigmoid(logits.detach()), m.cuda()).item()
    assert abs(d_ref - d_dev) < 1e-3
    i_ref = O.mean_iou(torch.sigmoid(ref_logits.detach()), m).item()
    i_dev = B.metrics.MeanIoU()(torch.sigmoid(logits.detach()), m.cuda()).item()
    assert abs(i_ref - i_dev) < 1e-3
    g = _grads(net)
    keys = sorted(ref_g)
    # headline: concatenated parameter gradient within 2e-2 of the fp32 oracle
    assert _global_rel(g, ref_g, keys) < 2e-2, _global_rel(g, ref_g, keys)
    gnorm = torch.cat([ref_g[k].flatten() for k in keys]).norm().item()
    for k in keys:  # per tensor: 1e-1 relative with an absolute floor tied to the global norm
        err = (g[k] - ref_g[k]).norm().item()
        assert err < 0.1 * ref_g[k].norm().item() + 5e-3 * gnorm, (k, err, ref_g[k].norm().item())


def test_unet2d_trained_parity(B):
    torch.manual_seed(0)
    model = B.UNetModel(B.UNetConfig(1, 1))
    sd = {k[len("unet."):]: v.clone() for k, v in model.state_dict().items()}
    shape = (2, 1, 32, 32)
    bce = lambda lo, m: F.binary_cross_entropy_with_logits(lo, m)  # noqa: E731
    sd = _train_oracle(sd, O.unet2d_forward, bce, shape, steps=12, lr=2e-4)
    x, m = _blob_task((2, 1, 64, 64), 999)
    for v in sd.values():
        v.grad = None
    ref_logits = O.unet2d_forward(sd, x)
    ref_loss = bce(ref_logits, m)
    ref_loss.backward()
    ref_g = {k: v.grad for k, v in sd.items()}
    model.load_state_dict({"unet." + k: v.detach() for k, v in sd.items()})
    model = model.cuda().train()
    out = model(images=x.cuda(), labels=m.cuda())
    out.loss.backward()
    lg = out.logits.detach().cpu()
    assert rel_l2(lg, ref_logits.detach()) < 2e-2
    agree = ((lg > 0) == (ref_logits.detach() > 0)).float().mean().item()
    assert agree >= 0.999, agree
    assert abs(out.loss.item() - ref_loss.item()) < 1e-3
    g = {k[len("unet."):]: v for k, v in _grads(model).items()}
    assert _global_rel(g, ref_g, sorted(ref_g)) < 2e-2
    # metrics through the reference-facing API (numpy in, floats out)
    mm = B.metrics.compute_metrics((lg.numpy(), m.numpy()))
    rr = OM.compute_metrics(ref_logits.detach().numpy(), m.numpy())
    assert abs(mm["iou"] - rr["iou"]) < 1e-3 and abs(mm["dice"] - rr["dice"]) < 1e-3


# ----------------------------------------------------------------------------------- losses / metrics
def test_losses_vs_reference_golden(B):
    f = load_golden("losses_metrics")
    t = torch.from_numpy(f["t3"]).cuda()
    cases = {
        "dice_sigmoid": B.DiceLoss(),
        "dice_softmax": B.DiceLoss(normalization="softmax"),
        "dice_weighted": B.DiceLoss(weight=torch.tensor([0.2, 0.3, 0.5])).cuda(),
        "bce_dice": B.BCEDiceLoss(1.0, 1.0),
        "bce_dice_a": B.BCEDiceLoss(0.3, 0.7),
    }
    for nm, crit in cases.items():
        x = torch.from_numpy(f["x3"]).cuda().requires_grad_(True)
        l = crit(x, t)
        l.backward()
        assert abs(l.item() - float(f[nm + "_loss"])) < 1e-5, nm          # fp32 inputs: float round-off
        assert rel_l2(x.grad.cpu(), f[nm + "_grad"]) < 1e-4, nm
    x = torch.from_numpy(f["x3"]).cuda()
    assert abs(B.DiceLoss()(x, torch.zeros_like(t)).item() - float(f["dice_zero_target_loss"])) < 1e-5
    pcd = B.compute_per_channel_dice(torch.sigmoid(x), t).cpu().numpy()
    np.testing.assert_allclose(pcd, f["per_channel_dice"], rtol=1e-5)
    t2 = torch.from_numpy(f["t2"]).cuda()
    for nm, crit in {"iou": B.IoULoss(), "f1": B.F1Loss(), "bce": B.BCEWithLogitsLoss()}.items():
        x2 = torch.from_numpy(f["x2"]).cuda().requires_grad_(True)
        l = crit(x2, t2)
        l.backward()
        assert abs(l.item() - float(f[nm + "_loss"])) < 1e-5, nm
        assert rel_l2(x2.grad.cpu(), f[nm + "_grad"]) < 1e-4, nm
    # SegmentationLoss (loss.py:59-70) without its third-party MS-SSIM term = F1 + IoU, one pass
    x2 = torch.from_numpy(f["x2"]).cuda().requires_grad_(True)
    l = B.SegmentationLoss(ms_ssim=False)(x2, t2)
    l.backward()
    assert abs(l.item() - float(f["iou_loss"]) - float(f["f1_loss"])) < 2e-5
    assert rel_l2(x2.grad.cpu(), f["iou_grad"] + f["f1_grad"]) < 1e-4
    assert B.flatten(x).shape == (x.shape[1], x.numel() // x.shape[1])
    assert torch.equal(B.flatten(x)[1], x[:, 1].reshape(-1))
    xc = torch.from_numpy(f["ce_x"]).cuda().requires_grad_(True)
    l = B.CrossEntropyLoss()(xc, torch.from_numpy(f["ce_y"]).cuda())
    l.backward()
    assert abs(l.item() - float(f["ce_loss"])) < 1e-5
    assert rel_l2(xc.grad.cpu(), f["ce_grad"]) < 1e-4


def test_metrics_vs_reference_golden(B):
    f = load_golden("losses_metrics")
    M = B.metrics
    logits, labels = f["x2"], f["t2"]
    m = M.compute_metrics((logits, labels))
    assert abs(m["iou"] - float(f["cm_iou"])) < 1e-6 and abs(m["dice"] - float(f["cm_dice"])) < 1e-6
    probs = 1 / (1 + np.exp(-logits[:, 0]))
    assert abs(M.compute_iou(probs, labels[:, 0]) - float(f["ci_05"])) < 1e-6
    assert abs(M.compute_dice(probs, labels[:, 0]) - float(f["cd_05"])) < 1e-6
    z = np.zeros_like(probs)
    assert M.compute_iou(z, z) == 0.0 and abs(M.compute_dice(z, z) - 0.5) < 1e-6
    iou, dice = M.iou_score(torch.from_numpy(logits), torch.from_numpy(labels))
    assert abs(iou - float(f["iou_score"])) < 1e-6 and abs(dice - float(f["iou_score_dice"])) < 1e-6
    assert abs(M.dice_coef(torch.from_numpy(logits), torch.from_numpy(labels)) - float(f["dice_coef"])) < 1e-5
    x3, t3, oh = (torch.from_numpy(f[k]).cuda() for k in ("x3", "t3", "oh5"))
    p5 = torch.softmax(x3, dim=1)
    assert abs(M.MeanIoU()(p5, oh).item() - float(f["miou_multi"])) < 1e-6
    assert abs(M.MeanIoU()(torch.sigmoid(x3[:, :1]), t3[:, :1]).item() - float(f["miou_single"])) < 1e-6
    assert abs(M.DiceCoefficient()(p5, oh).item() - float(f["dice_coeff"])) < 1e-5


# ----------------------------------------------------------------------------------- kernels, edge cases
def test_tcgen05_and_direct_kernels_agree(B):
    """Same GEMM problem through the tensor-core and the CUDA-core kernel: identical up to fp32
    summation order (both round the result to bf16 once)."""
    from mdeical_image_segmentation_b200 import ops
    from mdeical_image_segmentation_b200.ops import CL
    g = torch.Generator().manual_seed(0)
    for (n, d, h, w, cin, cout) in [(1, 1, 19, 23, 64, 64), (1, 6, 9, 10, 32, 48), (2, 4, 4, 4, 128, 256)]:
        x = CL(torch.randn((n, d, h, w, cin), generator=g).to(torch.bfloat16).cuda())
        dy = CL(torch.randn((n, d, h, w, cout), generator=g).to(torch.bfloat16).cuda())
        ks = (3, 3, 3) if d > 1 else (3, 3)
        wt = (torch.randn((cout, cin) + ks, generator=g) / (cin * 9) ** 0.5).cuda()
        fwd, bwd, k = ops.conv_repack_weight(wt)
        y1 = ops.conv_fwd(x, fwd, k, None, True, impl=1).dense().float()
        y2 = ops.conv_fwd(x, fwd, k, None, True, impl=2).dense().float()
        assert rel_l2(y1.cpu(), y2.cpu()) < 3e-3
        d1 = ops.conv_bwd_data(dy, bwd, k, impl=1).dense().float()
        d2 = ops.conv_bwd_data(dy, bwd, k, impl=2).dense().float()
        assert rel_l2(d1.cpu(), d2.cpu()) < 3e-3
        w1 = ops.conv_bwd_weight(x, dy, k, wt.shape, impl=1)
        w2 = ops.conv_bwd_weight(x, dy, k, wt.shape, impl=2)
        assert rel_l2(w1.cpu(), w2.cpu()) < 1e-4


def test_wide_weight_gradient_and_fused_bias_gradient(B):
    """Shapes that take the N = 128 four-tap-pair weight-gradient kernel + centre-tap GEMM (C_in >= 256,
    or C_in = 128 on a large plane), including a ragged plane and a channel slice of a wider buffer,
    against the CUDA-core kernel; and the bias gradient that rides in the same kernel against the
    stand-alone column-sum pass."""
    from mdeical_image_segmentation_b200 import ops
    from mdeical_image_segmentation_b200.ops import CL
    g = torch.Generator().manual_seed(11)
    cases = [(4, 1, 64, 64, 256, 128, 0), (2, 4, 24, 40, 256, 256, 0), (3, 1, 50, 44, 320, 192, 64),
             (16, 1, 128, 128, 128, 128, 0), (2, 1, 32, 24, 64, 64, 0), (1, 3, 16, 16, 32, 48, 0)]
    for (n, d, h, w, cin, cout, extra) in cases:
        xb = torch.randn((n, d, h, w, cin + extra), generator=g).to(torch.bfloat16).cuda()
        x = CL(xb, cin, extra)                          # channels [extra, extra + cin) of a wider buffer
        dy = CL(torch.randn((n, d, h, w, cout), generator=g).to(torch.bfloat16).cuda())
        ks = (3, 3, 3) if d > 1 else (3, 3)
        k = (3, 3, 3) if d > 1 else (1, 3, 3)
        wshape = (cout, cin) + ks
        w1, b1 = ops.conv_bwd_weight(x, dy, k, wshape, impl=1, bias_grad=True)
        w2 = ops.conv_bwd_weight(x, dy, k, wshape, impl=2)
        assert rel_l2(w1.cpu(), w2.cpu()) < 2e-4, (n, d, h, w, cin, cout)
        b2 = ops.bias_grad(dy)
        bref = dy.dense().float().sum(dim=(0, 1, 2, 3))
        assert rel_l2(b1.cpu(), bref.cpu()) < 1e-4 and rel_l2(b2.cpu(), bref.cpu()) < 1e-4


def test_maxpool_tie_routing_and_odd_sizes(B):
    """Post-ReLU all-zero windows: the gradient goes to the FIRST element in (d,h,w) order, as ATen."""
    from mdeical_image_segmentation_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.relu(torch.randn((2, 8, 5, 7, 9), generator=g) - 0.8)     # many zero windows, odd dims
    x = x.to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True)
    y = F.max_pool3d(xr, 2)
    gy = torch.randn(y.shape, generator=g).to(torch.bfloat16).float()
    y.backward(gy)
    xc = ops.from_nchw(x.cuda())
    yc = ops.maxpool_fwd(xc, 2)
    assert torch.equal(ops.to_nchw(yc, True).cpu(), y.detach())
    dx = ops.maxpool_bwd(xc, ops.from_nchw(gy.cuda()), 2)
    assert torch.equal(ops.to_nchw(dx, True).cpu(), xr.grad)


def test_nearest_resize_index_map(B):
    """F7: ConvTranspose output 2n-1 resized to 2n reads ct[max(i-1, 0)]; plus exact 2x."""
    from mdeical_image_segmentation_b200 import ops
    for (isz, osz) in [((3, 5, 7), (4, 6, 8)), ((4, 4, 4), (8, 8, 8)), ((15, 15, 15), (16, 16, 16))]:
        x = torch.arange(2 * 8 * isz[0] * isz[1] * isz[2], dtype=torch.float32).reshape((2, 8) + isz) % 251
        ref = F.interpolate(x, size=osz)
        out = ops.to_nchw(ops.upsample_nearest_fwd(ops.from_nchw(x.cuda()), osz), True).cpu()
        assert torch.equal(out, ref)
        xr = x.clone().requires_grad_(True)
        gy = (torch.arange(ref.numel(), dtype=torch.float32).reshape(ref.shape) % 7) - 3
        F.interpolate(xr, size=osz).backward(gy)
        dx = ops.to_nchw(ops.upsample_nearest_bwd(ops.from_nchw(gy.cuda()), isz), True).cpu()
        assert torch.equal(dx, xr.grad)


def test_deconv_plane_resize_backward_is_the_adjoint_of_forward(B):
    """Parity planes of ConvTranspose3d(k3,s2,p1) <-> nearest resize to the skip's size: the backward
    (exact-2x kernel with 8 loads in flight, and the generic kernel on a ragged size) must be the exact
    transpose of the forward gather — checked with small integers, so the identity is bit-exact."""
    from mdeical_image_segmentation_b200 import ops
    from mdeical_image_segmentation_b200.ops import CL
    g = torch.Generator().manual_seed(5)
    for (n, d, h, w, c, size) in [(2, 4, 6, 8, 16, (8, 12, 16)), (1, 3, 5, 4, 8, (7, 10, 9)),
                                  (1, 16, 16, 16, 64, (32, 32, 32))]:
        yp = CL(torch.randint(-3, 4, (8 * n, d, h, w, c), generator=g).to(torch.bfloat16).cuda())
        gy = CL(torch.randint(-3, 4, (n,) + size + (c,), generator=g).to(torch.bfloat16).cuda())
        y = ops.deconv_resize_fwd(yp, n, 3, size)
        dyp = ops.deconv_resize_bwd(gy, (n, d, h, w), 3)
        lhs = (y.dense().double() * gy.dense().double()).sum().item()
        rhs = (yp.dense().double() * dyp.dense().double()).sum().item()
        assert lhs == rhs, (n, d, h, w, size, lhs, rhs)
        assert dyp.dense().double().sum().item() == gy.dense().double().sum().item()   # every child lands once


def test_size_independent_properties_at_headline_size(B):
    """Linearity of the conv kernels and loss identities at a BASELINE-scale layer (128^3, 64 ch)."""
    from mdeical_image_segmentation_b200 import ops
    from mdeical_image_segmentation_b200.ops import CL
    torch.manual_seed(0)
    x = CL((torch.randn((1, 64, 128, 128, 32), device="cuda") * 0.5).to(torch.bfloat16))
    wt = torch.randn((32, 32, 3, 3, 3), device="cuda") * 0.03
    fwd, bwd, k = ops.conv_repack_weight(wt)
    y = ops.conv_fwd(x, fwd, k, None, False).dense().float()
    x2 = CL((x.dense().float() * 2).to(torch.bfloat16).contiguous())
    y2 = ops.conv_fwd(x2, fwd, k, None, False).dense().float()
    assert rel_l2((y2 / 2).cpu(), y.cpu()) < 4e-3                       # conv(2x) = 2 conv(x)
    # <conv(x), dy> == <x, dgrad(dy)> == <w, wgrad(x, dy)>  (adjoint identities)
    dy = CL((torch.randn_like(y) * 0.1).to(torch.bfloat16))
    dx = ops.conv_bwd_data(dy, bwd, k).dense().float()
    dw = ops.conv_bwd_weight(x, dy, k, wt.shape)
    lhs = (y.double() * dy.dense().double()).sum().item()
    mid = (x.dense().double() * dx.double()).sum().item()
    rhs = (wt.to(torch.bfloat16).double() * dw.double()).sum().item()
    assert abs(lhs - mid) < 2e-2 * abs(lhs) and abs(lhs - rhs) < 2e-2 * abs(lhs), (lhs, mid, rhs)
    m = (torch.rand((2, 1, 128, 128, 128), device="cuda") > 0.7).float()
    big = torch.where(m > 0, 30.0, -30.0)
    assert B.DiceLoss()(big, m).item() < 1e-5
    assert abs(B.metrics.MeanIoU()(m.clone(), m.clone()).item() - 1.0) < 1e-6


def test_empty_batch_is_a_noop(B):
    from mdeical_image_segmentation_b200 import ops
    x = ops.CL(torch.empty((0, 1, 8, 8, 16), dtype=torch.bfloat16, device="cuda"))
    y = ops.maxpool_fwd(x, 1)
    assert y.n == 0


def test_fused_epilogue_statistics(B):
    """Column statistics fused into the GEMM epilogues (b200_conv_*_stats) equal the sums of the tensor
    the kernel stored: (sum y, sum y^2) in forward; (sum dx, sum dx*dot) with the ReLU mask applied in
    backward-data; tensor-core and CUDA-core paths agree; channel-slice outputs and ragged planes."""
    from mdeical_image_segmentation_b200 import ops
    from mdeical_image_segmentation_b200.ops import CL
    g = torch.Generator().manual_seed(3)
    cases = [(2, 1, 40, 24, 64, 64), (2, 6, 20, 12, 32, 96), (3, 4, 4, 4, 128, 256), (1, 1, 19, 23, 64, 16),
             (2, 1, 64, 64, 128, 512)]
    for (n, d, h, w, cin, cout) in cases:
        x = CL(torch.randn((n, d, h, w, cin), generator=g).to(torch.bfloat16).cuda())
        dy = CL(torch.randn((n, d, h, w, cout), generator=g).to(torch.bfloat16).cuda())
        ks = (3, 3, 3) if d > 1 else (3, 3)
        wt = (torch.randn((cout, cin) + ks, generator=g) / (cin * 9) ** 0.5).cuda()
        bias = torch.randn(cout, generator=g).cuda()
        fwd, bwd, k = ops.conv_repack_weight(wt)
        for impl in (1, 2):
            # forward into a channel slice of a wider buffer
            cat = CL.empty(n, d, h, w, cout + 32, "cuda")
            y, cs = ops.conv_fwd(x, fwd, k, bias, True, out=cat.slice(32, cout), impl=impl, stats=True)
            yd = y.dense().double()
            ref = torch.stack([yd.sum((1, 2, 3)), (yd * yd).sum((1, 2, 3))], 1)      # [n][2][c]
            assert rel_l2(cs.cpu(), ref.cpu()) < 2e-5, (impl, n, d, h, w, cin, cout)
            # plain call stores the same tensor
            y0 = ops.conv_fwd(x, fwd, k, bias, True, impl=impl).dense()
            assert torch.equal(y0, y.dense())
            # backward data with mask + dot
            dx, cs2 = ops.conv_bwd_data(dy, bwd, k, mask=x, impl=impl, stats=True, dot=x)
            dxd = dx.dense().double()
            assert bool((dxd[x.dense() <= 0] == 0).all())
            ref2 = torch.stack([dxd.sum((1, 2, 3)), (dxd * x.dense().double()).sum((1, 2, 3))], 1)
            assert rel_l2(cs2.cpu(), ref2.cpu()) < 2e-5, (impl, "bwd")
            dx0 = ops.conv_bwd_data(dy, bwd, k, mask=x, impl=impl).dense()
            assert torch.equal(dx0, dx.dense())
            # sum-only and sum + sum-of-squares variants, without a mask
            dx1, cs3 = ops.conv_bwd_data(dy, bwd, k, impl=impl, stats="sum")
            d1 = dx1.dense().double()
            assert rel_l2(cs3[:, 0].cpu(), d1.sum((1, 2, 3)).cpu()) < 2e-5 and float(cs3[:, 1].abs().max()) == 0.0
            dx2, cs4 = ops.conv_bwd_data(dy, bwd, k, impl=impl, stats=True)
            ref4 = torch.stack([d1.sum((1, 2, 3)), (d1 * d1).sum((1, 2, 3))], 1)
            assert rel_l2(cs4.cpu(), ref4.cpu()) < 2e-5 and torch.equal(dx1.dense(), dx2.dense())
    # ConvTranspose 2x2 backward-data with the fused column sum
    xt = CL(torch.randn((2, 1, 16, 24, 128), generator=g).to(torch.bfloat16).cuda())
    dyt = CL(torch.randn((2, 1, 32, 48, 64), generator=g).to(torch.bfloat16).cuda())
    wtt = (torch.randn((128, 64, 2, 2), generator=g) / 16).cuda()
    wf, wb = ops.convt_repack_weight(wtt)
    for impl in (1, 2):
        dxt, cst = ops.convt_bwd_data(dyt, wb, mask=xt, impl=impl, stats=True)
        dd = dxt.dense().double()
        reft = torch.stack([dd.sum((1, 2, 3)), (dd * dd).sum((1, 2, 3))], 1)
        assert rel_l2(cst.cpu(), reft.cpu()) < 2e-5
        # scatter epilogue (forward) against the CUDA-core kernel
    y1 = ops.convt_fwd(xt, wf, None, impl=1).dense().float()
    y2 = ops.convt_fwd(xt, wf, None, impl=2).dense().float()
    assert rel_l2(y1.cpu(), y2.cpu()) < 3e-3
    # planes large enough for the weight-resident schedules (one column tile per CTA; 1 and 2 column tiles)
    for (n, hh, ww, ci, co) in [(4, 128, 128, 128, 64), (2, 128, 128, 256, 128)]:
        xb = CL(torch.randn((n, 1, hh, ww, ci), generator=g).to(torch.bfloat16).cuda())
        wb_ = (torch.randn((ci, co, 2, 2), generator=g) / ci ** 0.5).cuda()
        bb = torch.randn(co, generator=g).cuda()
        wf2, wb2 = ops.convt_repack_weight(wb_)
        a1 = ops.convt_fwd(xb, wf2, bb, impl=1).dense().float()
        a2 = ops.convt_fwd(xb, wf2, bb, impl=2).dense().float()
        assert rel_l2(a1.cpu(), a2.cpu()) < 3e-3, (n, hh, ww, ci, co)
        dyb = CL(torch.randn((n, 1, 2 * hh, 2 * ww, co), generator=g).to(torch.bfloat16).cuda())
        b1 = ops.convt_bwd_data(dyb, wb2, mask=xb, impl=1).dense().float()
        b2 = ops.convt_bwd_data(dyb, wb2, mask=xb, impl=2).dense().float()
        assert rel_l2(b1.cpu(), b2.cpu()) < 3e-3, (n, hh, ww, ci, co)
    # weight-resident halo schedule (2D, C_in * C_out small) against the streaming one
    lib = __import__("mdeical_image_segmentation_b200._lib", fromlist=["lib"]).lib()
    xr = CL(torch.randn((4, 1, 96, 80, 64), generator=g).to(torch.bfloat16).cuda())
    for cout in (64, 128):
        wr = (torch.randn((cout, 64, 3, 3), generator=g) / 24).cuda()
        fr, br, kr = ops.conv_repack_weight(wr)
        r1 = ops.conv_fwd(xr, fr, kr, None, True, impl=1).dense()
        lib.b200_set_tuning(3, 0)
        r0 = ops.conv_fwd(xr, fr, kr, None, True, impl=1).dense()
        lib.b200_set_tuning(3, 1)
        assert torch.equal(r0, r1), cout
        for ew in (4, 8):   # 4 and 8 epilogue warps store the same tensor and the same statistics
            lib.b200_set_tuning(5, ew)
            r2, s2 = ops.conv_fwd(xr, fr, kr, None, True, impl=1, stats=True)
            lib.b200_set_tuning(5, 0)
            assert torch.equal(r2.dense(), r1), (cout, ew)
            rd = r1.double()
            refs = torch.stack([rd.sum((1, 2, 3)), (rd * rd).sum((1, 2, 3))], 1)
            assert rel_l2(s2.cpu(), refs.cpu()) < 2e-5, (cout, ew)
