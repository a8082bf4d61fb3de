"""Pins the CPU oracle (oracle/) against the golden vectors produced by the reference itself
(oracle/make_golden.py).  fp32 on both sides: agreement to float round-off."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2, sd_from
from oracle import metrics as OM
from oracle import unet as O

TOL = 2e-5


def _seeded_unet2d_sd(cin, cout, seed):
    # the 2D fixture stores the init seed: rebuild the weights with the product module's constructor
    # (parameter-creation order identical to the reference; checked by make_golden.check_seed_init)
    from mdeical_image_segmentation_b200.unet2d import UNet
    torch.manual_seed(seed)
    return {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in UNet(cin, cout).state_dict().items()}


@pytest.mark.parametrize("name,cin,cout", [("unet2d_1_1", 1, 1), ("unet2d_3_4", 3, 4)])
def test_unet2d_matches_reference(name, cin, cout):
    fix = load_golden(name)
    sd = _seeded_unet2d_sd(cin, cout, int(fix["seed"]))
    x = torch.from_numpy(fix["x"])
    y = torch.from_numpy(fix["y"])
    logits = O.unet2d_forward(sd, x)
    assert rel_l2(logits, fix["logits"]) < TOL
    loss = O.unet_model_loss(logits, y, cout)
    assert abs(loss.item() - float(fix["loss"])) < 1e-5
    loss.backward()
    names = [str(s) for s in fix["grad_names"]]
    l2 = np.array([sd[k].grad.norm().item() for k in names])
    np.testing.assert_allclose(l2, fix["grad_l2"], rtol=2e-3, atol=1e-7)
    for k, v in fix.items():
        if k.startswith("grad__"):
            assert rel_l2(sd[k[6:]].grad, v) < 1e-3, k


@pytest.mark.parametrize("name", ["unet3d_gcr_nearest", "unet3d_cbr_deconv", "unet3d_cr_nearest",
                                  "unet2d3_gcr_nearest", "unet2d3_cbr_deconv"])
def test_unet3d_family_matches_reference(name):
    fix = load_golden(name)
    sd = sd_from(fix)
    for v in sd.values():
        if v.is_floating_point() and v.dim() > 0:
            v.requires_grad_(True)
    # running stats / counters are buffers
    for k in sd:
        if "running_" in k or "num_batches" in k:
            sd[k] = sd[k].detach()
    kw = dict(num_levels=3, layer_order=str(fix["layer_order"]), num_groups=4,
              upsample=str(fix["upsample"]), is3d=bool(fix["is3d"]))
    x, t = torch.from_numpy(fix["x"]), torch.from_numpy(fix["t"])
    logits = O.unet3d_forward(sd, x, training=True, **kw)
    assert rel_l2(logits, fix["logits"]) < TOL
    loss = O.bce_dice_loss(logits, t)
    assert abs(loss.item() - float(fix["loss"])) < 1e-5
    loss.backward()
    for k, v in fix.items():
        if k.startswith("grad__"):
            g = sd[k[6:]].grad
            err = (g - torch.from_numpy(v)).norm().item()
            assert err < 1e-3 * (np.linalg.norm(v) + 1e-4), (k, err)
        if k.startswith("after__"):
            assert rel_l2(sd[k[7:]], v) < 1e-5, k
    with torch.no_grad():
        ev = O.unet3d_forward(sd, x, training=False, **kw)
    assert rel_l2(ev, fix["logits_eval"]) < TOL


def test_unetconv2_block():
    fix = load_golden("unetconv2_bn")
    sd = sd_from(fix)
    x = torch.from_numpy(fix["x"]).requires_grad_(True)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    y = O.unet_conv2(sd, "", x, n=2, training=True)
    assert rel_l2(y, fix["y"]) < TOL
    y.square().sum().backward()
    assert rel_l2(x.grad, fix["dx"]) < 1e-4


def test_losses():
    f = load_golden("losses_metrics")
    t = torch.from_numpy(f["t3"])
    cases = {
        "dice_sigmoid": lambda x: O.dice_loss(x, t),
        "dice_softmax": lambda x: O.dice_loss(x, t, normalization="softmax"),
        "dice_weighted": lambda x: O.dice_loss(x, t, weight=torch.tensor([0.2, 0.3, 0.5])),
        "bce_dice": lambda x: O.bce_dice_loss(x, t),
        "bce_dice_a": lambda x: O.bce_dice_loss(x, t, 0.3, 0.7),
    }
    for nm, fn in cases.items():
        x = torch.from_numpy(f["x3"]).requires_grad_(True)
        l = fn(x)
        l.backward()
        assert abs(l.item() - float(f[nm + "_loss"])) < 1e-6, nm
        assert rel_l2(x.grad, f[nm + "_grad"]) < 1e-5, nm
    x = torch.from_numpy(f["x3"])
    assert abs(O.dice_loss(x, torch.zeros_like(t)).item() - float(f["dice_zero_target_loss"])) < 1e-6
    np.testing.assert_allclose(O.per_channel_dice(torch.sigmoid(x), t).numpy(), f["per_channel_dice"], rtol=1e-6)
    t2 = torch.from_numpy(f["t2"])
    for nm, fn in {"iou": O.iou_loss, "f1": O.f1_loss,
                   "bce": lambda a, b: O.unet_model_loss(a, b, 1)}.items():
        x2 = torch.from_numpy(f["x2"]).requires_grad_(True)
        l = fn(x2, t2)
        l.backward()
        assert abs(l.item() - float(f[nm + "_loss"])) < 1e-6, nm
        assert rel_l2(x2.grad, f[nm + "_grad"]) < 1e-5, nm
    xc = torch.from_numpy(f["ce_x"]).requires_grad_(True)
    l = O.unet_model_loss(xc, torch.from_numpy(f["ce_y"]), 4)
    l.backward()
    assert abs(l.item() - float(f["ce_loss"])) < 1e-6
    assert rel_l2(xc.grad, f["ce_grad"]) < 1e-5


def test_metrics():
    f = load_golden("losses_metrics")
    logits, labels = f["x2"], f["t2"]
    m = OM.compute_metrics(logits, labels)
    assert abs(m["iou"] - float(f["cm_iou"])) < 1e-7 and abs(m["dice"] - float(f["cm_dice"])) < 1e-7
    probs = 1 / (1 + np.exp(-logits[:, 0]))
    assert abs(OM.compute_iou(probs, labels[:, 0]) - float(f["ci_05"])) < 1e-7
    assert abs(OM.compute_dice(probs, labels[:, 0]) - float(f["cd_05"])) < 1e-7
    z = np.zeros_like(probs)
    assert OM.compute_iou(z, z) == float(f["ci_empty"]) == 0.0
    assert abs(OM.compute_dice(z, z) - float(f["cd_empty"])) < 1e-9       # (0 + e) / (e + e) = 0.5
    iou, dice = OM.iou_score(logits, labels)
    assert abs(iou - float(f["iou_score"])) < 1e-9 and abs(dice - float(f["iou_score_dice"])) < 1e-9
    assert abs(OM.dice_coef(logits, labels) - float(f["dice_coef"])) < 1e-6
    x3, t3, oh = torch.from_numpy(f["x3"]), torch.from_numpy(f["t3"]), torch.from_numpy(f["oh5"])
    p5 = torch.softmax(x3, dim=1)
    assert abs(O.mean_iou(p5, oh).item() - float(f["miou_multi"])) < 1e-6
    assert abs(O.mean_iou(torch.sigmoid(x3[:, :1]), t3[:, :1]).item() - float(f["miou_single"])) < 1e-6
    assert abs(O.dice_coefficient(p5, oh).item() - float(f["dice_coeff"])) < 1e-6


def test_analytic_known_answers():
    m = torch.zeros(1, 1, 4, 4, 4)
    m[0, 0, 1:3] = 1
    big = torch.where(m > 0, torch.tensor(30.0), torch.tensor(-30.0))
    assert O.dice_loss(big, m).item() < 1e-6                      # identical masks -> dice 1
    assert abs(O.mean_iou(m.clone(), m.clone()).item() - 1.0) < 1e-7
    assert abs(O.dice_loss(-big, m).item() - 1.0) < 1e-6          # disjoint masks -> dice 0


def _unet3plus_sd(seed, requires_grad=True, deep_sup=False, cls_name=None):
    """UNet 3+ weights of the fixture: the product module's state_dict layout (checked against the
    reference's key list stored in the fixture) filled by oracle.unet3plus.fill_state_dict_(seed)."""
    from mdeical_image_segmentation_b200 import unet3plus as P3
    from oracle import unet3plus as O3
    cls = getattr(P3, cls_name) if cls_name else (P3.UNet_3Plus_DeepSup if deep_sup else P3.UNet_3Plus)
    sd = {k: v.clone() for k, v in cls(3, 1).state_dict().items()}
    O3.fill_state_dict_(sd, seed)
    if requires_grad:
        sd = {k: (v.requires_grad_(True) if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    return sd


def test_unet3plus_matches_reference():
    """UNet 3+ (model/unet2d/unet.py:136-446): state_dict keys, logits, loss, running statistics and
    gradients of the oracle against the reference's own run (fixture made by oracle/make_golden.py)."""
    from oracle import unet3plus as O3
    fix = load_golden("unet3plus_3_1")
    sd = _unet3plus_sd(int(fix["seed"]))
    assert list(sd.keys()) == [str(k) for k in fix["state_dict_keys"]]
    x, y = torch.from_numpy(fix["x"]), torch.from_numpy(fix["y"])
    logits = O3.unet3plus_forward(sd, x, training=True)
    assert rel_l2(logits, fix["logits"]) < TOL
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
    assert abs(loss.item() - float(fix["loss"])) < 1e-5
    assert rel_l2(sd["bn1d_1.running_mean"], fix["running_mean_bn1d_1"]) < 1e-4
    assert rel_l2(sd["conv1.conv1.1.running_var"], fix["running_var_conv1"]) < 1e-4
    loss.backward()
    names = [str(s) for s in fix["grad_names"]]
    l2 = np.array([sd[k].grad.norm().item() for k in names])
    np.testing.assert_allclose(l2, fix["grad_norms"], rtol=5e-3, atol=1e-6)
    for k, v in fix.items():
        if k.startswith("grad::"):
            assert rel_l2(sd[k[6:]].grad, v) < 2e-3, k


def test_unet3plus_deepsup_matches_reference():
    """UNet_3Plus_DeepSup (unet.py:454-787): the five per-scale logits, the summed BCE loss
    (unet.py:1203-1208) and the gradients of the oracle against the reference's own run."""
    from oracle import unet3plus as O3
    fix = load_golden("unet3plus_deepsup_3_1")
    sd = _unet3plus_sd(int(fix["seed"]), deep_sup=True)
    assert list(sd.keys()) == [str(k) for k in fix["state_dict_keys"]]
    x, y = torch.from_numpy(fix["x"]), torch.from_numpy(fix["y"])
    ds = O3.unet3plus_forward(sd, x, training=True, deep_sup=True)
    assert len(ds) == 5
    for i, d in enumerate(ds):
        assert rel_l2(d, fix[f"d{i + 1}"]) < TOL, i
    loss = sum(torch.nn.functional.binary_cross_entropy_with_logits(d, y) for d in ds)
    assert abs(loss.item() - float(fix["loss"])) < 1e-5
    loss.backward()
    names = [str(s) for s in fix["grad_names"]]
    l2 = np.array([sd[k].grad.norm().item() for k in names])
    np.testing.assert_allclose(l2, fix["grad_norms"], rtol=5e-3, atol=1e-6)
    for k, v in fix.items():
        if k.startswith("grad::"):
            assert rel_l2(sd[k[6:]].grad, v) < 2e-3, k


def test_msssim_oracle_known_answers():
    """oracle/msssim.py restates pytorch-msssim 1.0.0 (absent here: parity unpinned for this term).  Known
    answers: identical images -> 1; the separable valid Gaussian equals a dense 11 x 11 correlation; a single
    scale agrees with SSIM computed from its definition; the size assertion; symmetry in (x, y)."""
    from oracle import msssim as M
    g = torch.Generator().manual_seed(0)
    x = torch.rand((2, 1, 176, 192), generator=g)
    y = torch.rand((2, 1, 176, 192), generator=g)
    assert abs(M.ms_ssim(x, x).item() - 1.0) < 1e-6
    assert abs(M.ms_ssim(x, y).item() - M.ms_ssim(y, x).item()) < 1e-6
    win = M.fspecial_gauss_1d()
    assert abs(win.sum().item() - 1.0) < 1e-6 and win.shape == (1, 1, 11)
    dense = (win.view(11, 1) * win.view(1, 11)).view(1, 1, 11, 11)
    a = M.gaussian_filter(x, win)
    b = torch.nn.functional.conv2d(x, dense)
    assert a.shape == (2, 1, 166, 182) and (a - b).abs().max() < 1e-6
    mu1, mu2 = b, torch.nn.functional.conv2d(y, dense)
    s1 = torch.nn.functional.conv2d(x * x, dense) - mu1 ** 2
    s2 = torch.nn.functional.conv2d(y * y, dense) - mu2 ** 2
    s12 = torch.nn.functional.conv2d(x * y, dense) - mu1 * mu2
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ssim = ((2 * mu1 * mu2 + c1) * (2 * s12 + c2)) / ((mu1 ** 2 + mu2 ** 2 + c1) * (s1 + s2 + c2))
    got, _ = M.ssim_maps(x, y)
    assert (got.flatten() - ssim.flatten(1).mean(1)).abs().max() < 1e-5
    import pytest
    with pytest.raises(AssertionError):
        M.ms_ssim(x[..., :160, :], y[..., :160, :])
    t = (x > 0.5).float()
    assert M.msssim_loss(40 * (t - 0.5), t).item() < 1e-6


def test_msssim_oracle_vs_independent_float64_numpy():
    """A second, independent restatement (numpy / scipy, float64, dense loops over the five scales) of the published
    multi-scale SSIM with pytorch-msssim's conventions — valid 11-tap Gaussian (sigma 1.5), K = (0.01, 0.03), 2 x 2
    average pooling with zero padding of odd sizes counted in the divisor, cs of scales 0-3 and ssim of scale 4 raised
    to the weights — must agree with oracle/msssim.py on odd and even image sizes."""
    import numpy as np
    from scipy.ndimage import correlate1d
    from oracle import msssim as M

    def blur(a):
        k = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))
        k /= k.sum()
        a = correlate1d(a, k, axis=0, mode="constant")[5:-5]
        return correlate1d(a, k, axis=1, mode="constant")[:, 5:-5]

    def pool(a):
        h, w = a.shape
        a = np.pad(a, ((h % 2, h % 2), (w % 2, w % 2)))              # F.avg_pool2d(padding = size % 2), zeros counted
        h2, w2 = a.shape[0] // 2, a.shape[1] // 2
        return a[:2 * h2, :2 * w2].reshape(h2, 2, w2, 2).mean(axis=(1, 3))

    def ms_ssim(x, y):
        vals = []
        for lvl in range(5):
            mu1, mu2 = blur(x), blur(y)
            s1, s2, s12 = blur(x * x) - mu1 ** 2, blur(y * y) - mu2 ** 2, blur(x * y) - mu1 * mu2
            cs = (2 * s12 + 0.03 ** 2) / (s1 + s2 + 0.03 ** 2)
            ssim = (2 * mu1 * mu2 + 0.01 ** 2) / (mu1 ** 2 + mu2 ** 2 + 0.01 ** 2) * cs
            vals.append(max(cs.mean(), 0.0) if lvl < 4 else max(ssim.mean(), 0.0))
            x, y = pool(x), pool(y)
        return float(np.prod(np.array(vals) ** np.array(M.WEIGHTS)))

    rng = np.random.default_rng(3)
    for shape in ((176, 192), (177, 201), (256, 163)):
        x = rng.random(shape)
        y = np.clip(x + 0.3 * rng.standard_normal(shape), 0, 1)
        want = ms_ssim(x, y)
        got = M.ms_ssim(torch.from_numpy(x)[None, None], torch.from_numpy(y)[None, None], size_average=True).item()
        assert abs(got - want) < 5e-6, (shape, got, want)      # the package (and the oracle) build the window in fp32


def test_unet3plus_cgm_matches_reference():
    """UNet_3Plus_DeepSup_CGM (unet.py:795-1153, classification dropout off): gated probabilities of the five
    scales, BCE loss, gradient norms (the cls convolution receives none: arg-max) against the reference's run."""
    from oracle import unet3plus as O3
    fix = load_golden("unet3plus_cgm_3_1")
    sd = _unet3plus_sd(int(fix["seed"]), cls_name="UNet_3Plus_DeepSup_CGM")
    assert list(sd.keys()) == [str(k) for k in fix["state_dict_keys"]]
    x, y = torch.from_numpy(fix["x"]), torch.from_numpy(fix["y"])
    ds = O3.unet3plus_forward(sd, x, training=True, deep_sup=True, cgm=True)
    gates = fix["cls"].argmax(1)
    assert 0 < gates.sum() < len(gates)                      # the fixture exercises both gate values
    for i, d in enumerate(ds):
        assert (d.detach() - torch.from_numpy(fix[f"d{i + 1}"])).abs().max() < 1e-5, i
        assert torch.all(d.detach()[gates == 0] == 0.5)
    loss = sum(torch.nn.functional.binary_cross_entropy(d, y) for d in ds)
    assert abs(loss.item() - float(fix["loss"])) < 1e-5
    loss.backward()
    names = [str(s) for s in fix["grad_names"]]
    for k, has, norm in zip(names, fix["has_grad"], fix["grad_norms"]):
        if not has:
            assert k.startswith("cls.") and sd[k].grad is None
        else:
            assert abs(sd[k].grad.norm().item() - norm) < 5e-3 * norm + 1e-6, k


def test_predictor_patch_logic_matches_reference():
    """SliceBuilder._build_slices / mirror_pad / remove_padding of the reference (fixture made from
    dataset/unet3d_dataset/utils.py) against the oracle restatement AND the product's host-side logic."""
    from oracle import predictor as OP
    from mdeical_image_segmentation_b200 import predictor as PP
    fix = load_golden("predictor_slices")
    for i, (shape, patch, stride) in enumerate(fix["cases"].tolist()):
        want = fix[f"slices{i}"].tolist()
        for build in (OP.build_slices, PP.build_slices):
            got = [[s.start for s in idx] + [s.stop for s in idx] for idx in build(shape, patch, stride)]
            assert got == want, (i, build.__module__)
    assert np.array_equal(OP.mirror_pad(fix["pad_in3"], (2, 3, 4)), fix["pad_out3"])
    assert np.array_equal(OP.mirror_pad(fix["pad_in4"], (1, 0, 5)), fix["pad_out4"])
    assert np.array_equal(OP.remove_padding(fix["pad_out4"][None], (1, 0, 5))[0], fix["unpad_out4"])
    assert np.array_equal(fix["unpad_out4"], fix["pad_in4"])
    assert PP.gen_indices(10, 10, 3) == [0] and PP.gen_indices(11, 4, 4) == [0, 4, 7]
    import pytest
    with pytest.raises(AssertionError):
        PP.gen_indices(3, 4, 1)
    # the whole oracle loop: an identity model through overlapping patches and a halo gives the volume back
    rng = np.random.default_rng(1)
    raw = rng.standard_normal((20, 70, 66)).astype(np.float32)
    out = OP.standard_predict(lambda b: b, raw, (8, 64, 64), (5, 3, 2), halo_shape=(2, 4, 4))
    assert out.shape == (1, 20, 70, 66) and np.abs(out[0] - raw).max() < 1e-6


@pytest.mark.parametrize("name,cls", [("resunet3d_gcr", "ResidualUNet3D"), ("resunetse3d_gcr", "ResidualUNetSE3D")])
def test_residual_unet3d_matches_reference(name, cls):
    """ResidualUNet3D / ResidualUNetSE3D (model.py:197-280): oracle restatement against the reference's run —
    logits, loss, every parameter gradient — and the product module's state_dict layout against the reference's."""
    fix = load_golden(name)
    sd = sd_from(fix)
    for v in sd.values():
        v.requires_grad_(True)
    x, t = torch.from_numpy(fix["x"]), torch.from_numpy(fix["t"])
    logits = O.residual_unet3d_forward(sd, x, num_levels=3, num_groups=4)
    assert rel_l2(logits, fix["logits"]) < TOL
    loss = O.bce_dice_loss(logits, t)
    assert abs(loss.item() - float(fix["loss"])) < 1e-5
    loss.backward()
    for k, v in fix.items():
        if k.startswith("grad__"):
            err = (sd[k[6:]].grad - torch.from_numpy(v)).norm().item()
            assert err < 1e-3 * (np.linalg.norm(v) + 1e-4), (k, err)
    import mdeical_image_segmentation_b200 as B
    net = getattr(B, cls)(2, 2, f_maps=[8, 16, 32], num_groups=4)
    assert list(net.state_dict().keys()) == [k[4:] for k in fix if k.startswith("sd__")]
    net.load_state_dict(sd_from(fix), strict=True)
    assert [tuple(v.shape) for v in net.state_dict().values()] == [fix[k].shape for k in fix if k.startswith("sd__")]
    default = getattr(B, cls)(1, 1)                      # 5 levels, 64..1024 feature maps (model.py:207-213)
    assert len(default.encoders) == 5 and len(default.decoders) == 4
    assert default.encoders[4].basic_module.conv2.conv.out_channels == 1024
    elu = B.ResidualUNet3D(1, 1, layer_order="cge", f_maps=[8, 16])      # round 2: ELU / LeakyReLU orders build
    assert isinstance(elu.encoders[0].basic_module.non_linearity, torch.nn.ELU)
    with pytest.raises(NotImplementedError):                              # the fused SE gate needs a ReLU output
        B.ResidualUNetSE3D(1, 1, layer_order="cge", f_maps=[8, 16])
