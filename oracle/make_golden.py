"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference via
oracle/ref_loader.py) on seeded inputs.  Run inside the build container:

    python -m oracle.make_golden

The fixtures pin the oracle (tests/test_oracle_golden.py); they are small on purpose: the full-width
2D ``UNet`` (31 M parameters) is stored as its init seed, not its weights — the reference, the oracle
and the product modules all create their ``nn.Conv2d`` children in the same order, so
``torch.manual_seed(seed)`` reproduces the same ``state_dict`` (checked here).
TEST INFRASTRUCTURE (see oracle/__init__).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def blob_target(shape, seed, q=0.7):
    """Coherent binary blobs: box-blurred noise thresholded at its q-quantile."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(shape, generator=g)
    nd = len(shape) - 2
    pool = torch.nn.functional.avg_pool3d if nd == 3 else torch.nn.functional.avg_pool2d
    blur = pool(noise, 5, stride=1, padding=2)
    thr = torch.quantile(blur.flatten(), q)
    return (blur > thr).float()


def gen_unet2d(ns):
    for name, cin, cout in (("unet2d_1_1", 1, 1), ("unet2d_3_4", 3, 4)):
        torch.manual_seed(0)
        net = ns.U2.UNet(cin, cout)
        g = torch.Generator().manual_seed(1234)
        x = torch.randn((2, cin, 32, 32), generator=g)
        if cout == 1:
            y = (torch.rand((2, 1, 32, 32), generator=g) > 0.7).float()
            crit = torch.nn.BCEWithLogitsLoss()
        else:
            y = torch.randint(0, cout, (2, 32, 32), generator=g)
            crit = torch.nn.CrossEntropyLoss()
        logits = net(x)
        loss = crit(logits, y)
        loss.backward()
        grads = {k: p.grad for k, p in net.named_parameters()}
        keep = ["final_conv.weight", "final_conv.bias", "down_conv.0.first.weight",
                "down_conv.0.first.bias", "up_sample.3.up.bias", "up_sample.0.up.bias"]
        out = dict(seed=0, x=_np(x), y=_np(y), logits=_np(logits), loss=_np(loss),
                   grad_names=np.array(sorted(grads)),
                   grad_l2=np.array([grads[k].norm().item() for k in sorted(grads)]),
                   grad_sum=np.array([grads[k].double().sum().item() for k in sorted(grads)]))
        for k in keep:
            out["grad__" + k] = _np(grads[k])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "loss", loss.item())


def gen_unet3d(ns):
    cfgs = [
        ("unet3d_gcr_nearest", dict(layer_order="gcr", upsample="default"), True),
        ("unet3d_cbr_deconv", dict(layer_order="cbr", upsample="deconv"), True),
        ("unet3d_cr_nearest", dict(layer_order="cr", upsample="default"), True),
        ("unet2d3_gcr_nearest", dict(layer_order="gcr", upsample="default"), False),
        ("unet2d3_cbr_deconv", dict(layer_order="cbr", upsample="deconv"), False),
    ]
    for name, kw, is3d in cfgs:
        torch.manual_seed(0)
        cls = ns.M3.UNet3D if is3d else ns.M3.UNet2D
        cin, cout = (2, 2)
        net = cls(cin, cout, f_maps=[8, 16, 32], num_groups=4, **kw)
        net.train()
        shape = (2, cin, 16, 16, 16) if is3d else (2, cin, 32, 32)
        g = torch.Generator().manual_seed(77)
        x = torch.randn(shape, generator=g)
        t = blob_target((2, cout) + shape[2:], 5)
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        logits = net(x)
        loss = ns.L3.BCEDiceLoss(1.0, 1.0)(logits, t)
        loss.backward()
        out = dict(x=_np(x), t=_np(t), logits=_np(logits), loss=_np(loss), is3d=is3d,
                   layer_order=kw["layer_order"], upsample=kw["upsample"])
        for k, v in sd0.items():
            out["sd__" + k] = _np(v)
        for k, p in net.named_parameters():
            out["grad__" + k] = _np(p.grad)
        for k, v in net.state_dict().items():
            if "running_" in k:
                out["after__" + k] = _np(v)
        net.eval()
        with torch.no_grad():
            out["logits_eval"] = _np(net(x))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print(name, "loss", loss.item())


def gen_resunet3d(ns):
    """ResidualUNet3D (model/unet3d/model.py:197-238) from the reference: weights, logits, loss, gradients."""
    for name, cls in (("resunet3d_gcr", ns.M3.ResidualUNet3D), ("resunetse3d_gcr", ns.M3.ResidualUNetSE3D)):
        _gen_resunet3d(ns, name, cls)


def _gen_resunet3d(ns, name, cls):
    torch.manual_seed(0)
    net = cls(2, 2, f_maps=[8, 16, 32], num_groups=4)
    net.train()
    g = torch.Generator().manual_seed(78)
    x = torch.randn((2, 2, 16, 16, 16), generator=g)
    t = blob_target((2, 2, 16, 16, 16), 6)
    sd0 = {k: v.clone() for k, v in net.state_dict().items()}
    logits = net(x)
    loss = ns.L3.BCEDiceLoss(1.0, 1.0)(logits, t)
    loss.backward()
    out = dict(x=_np(x), t=_np(t), logits=_np(logits), loss=_np(loss))
    for k, v in sd0.items():
        out["sd__" + k] = _np(v)
    for k, p in net.named_parameters():
        out["grad__" + k] = _np(p.grad)
    net.eval()
    with torch.no_grad():
        out["logits_eval"] = _np(net(x))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item(), len(sd0), "tensors")


def gen_blocks(ns):
    torch.manual_seed(3)
    blk = ns.L2.unetConv2(3, 8, True)
    blk.train()
    g = torch.Generator().manual_seed(9)
    x = torch.randn((2, 3, 16, 16), generator=g, requires_grad=True)
    sd0 = {k: v.clone() for k, v in blk.state_dict().items()}
    y = blk(x)
    y.square().sum().backward()
    out = dict(x=_np(x), y=_np(y), dx=_np(x.grad))
    for k, v in sd0.items():
        out["sd__" + k] = _np(v)
    for k, p in blk.named_parameters():
        out["grad__" + k] = _np(p.grad)
    np.savez_compressed(os.path.join(OUT, "unetconv2_bn.npz"), **out)
    print("unetconv2 ok")


def gen_losses_metrics(ns):
    g = torch.Generator().manual_seed(11)
    out = {}
    # 3D dice-family losses on (N, C, D, H, W)
    x = torch.randn((2, 3, 6, 8, 10), generator=g, requires_grad=True)
    t = (torch.rand((2, 3, 6, 8, 10), generator=g) > 0.6).float()
    for nm, crit in (("dice_sigmoid", ns.L3.DiceLoss()),
                     ("dice_softmax", ns.L3.DiceLoss(normalization="softmax")),
                     ("dice_weighted", ns.L3.DiceLoss(weight=torch.tensor([0.2, 0.3, 0.5]))),
                     ("bce_dice", ns.L3.BCEDiceLoss(1.0, 1.0)),
                     ("bce_dice_a", ns.L3.BCEDiceLoss(0.3, 0.7))):
        x.grad = None
        l = crit(x, t)
        l.backward()
        out[nm + "_loss"] = _np(l)
        out[nm + "_grad"] = _np(x.grad)
    out["x3"], out["t3"] = _np(x), _np(t)
    out["dice_zero_target_loss"] = _np(ns.L3.DiceLoss()(x.detach(), torch.zeros_like(t)))
    out["per_channel_dice"] = _np(ns.L3.compute_per_channel_dice(torch.sigmoid(x.detach()), t))
    # 2D losses
    x2 = torch.randn((4, 1, 24, 24), generator=g, requires_grad=True)
    t2 = (torch.rand((4, 1, 24, 24), generator=g) > 0.7).float()
    for nm, crit in (("iou", ns.LOSS2.IoULoss()), ("f1", ns.LOSS2.F1Loss()),
                     ("bce", torch.nn.BCEWithLogitsLoss())):
        x2.grad = None
        l = crit(x2, t2)
        l.backward()
        out[nm + "_loss"] = _np(l)
        out[nm + "_grad"] = _np(x2.grad)
    out["x2"], out["t2"] = _np(x2), _np(t2)
    xc = torch.randn((2, 4, 12, 12), generator=g, requires_grad=True)
    yc = torch.randint(0, 4, (2, 12, 12), generator=g)
    l = torch.nn.CrossEntropyLoss()(xc, yc)
    l.backward()
    out["ce_x"], out["ce_y"], out["ce_loss"], out["ce_grad"] = _np(xc), _np(yc), _np(l), _np(xc.grad)
    # trainer/metrcis.py
    MX = ns.MX
    MX.plot_first_element = lambda *a, **k: None        # the matplotlib side effect is not on the path
    logits = _np(x2)
    labels = _np(t2)
    m = MX.compute_metrics((logits, labels)) if False else None
    from types import SimpleNamespace
    m = MX.compute_metrics(SimpleNamespace(predictions=logits, label_ids=labels))
    out["cm_iou"], out["cm_dice"] = np.float64(m["iou"]), np.float64(m["dice"])
    probs = 1 / (1 + np.exp(-logits[:, 0]))
    out["ci_05"] = np.float64(MX.compute_iou(probs, labels[:, 0], 0.5))
    out["cd_05"] = np.float64(MX.compute_dice(probs, labels[:, 0], 0.5))
    out["ci_empty"] = np.float64(MX.compute_iou(np.zeros_like(probs), np.zeros_like(probs), 0.5))
    out["cd_empty"] = np.float64(MX.compute_dice(np.zeros_like(probs), np.zeros_like(probs), 0.5))
    iou, dice = MX.iou_score(x2.detach(), t2)
    out["iou_score"], out["iou_score_dice"] = np.float64(iou), np.float64(dice)
    out["dice_coef"] = np.float64(MX.dice_coef(x2.detach(), t2))
    # model/unet3d/metrics.py
    p5 = torch.softmax(x.detach(), dim=1)
    oh = torch.nn.functional.one_hot(torch.randint(0, 3, (2, 6, 8, 10), generator=g), 3)
    oh = oh.permute(0, 4, 1, 2, 3).float()
    out["oh5"] = _np(oh)
    out["miou_multi"] = _np(ns.MET3.MeanIoU()(p5, oh.clone()))
    p1 = torch.sigmoid(x.detach()[:, :1])
    out["miou_single"] = _np(ns.MET3.MeanIoU()(p1, t[:, :1].clone()))
    out["dice_coeff"] = _np(ns.MET3.DiceCoefficient()(p5, oh))
    np.savez_compressed(os.path.join(OUT, "losses_metrics.npz"), **out)
    print("losses/metrics ok")


def check_seed_init(ns):
    """The 2D fixture stores a seed instead of 124 MB of weights: prove the three constructors agree."""
    import mdeical_image_segmentation_b200 as B
    torch.manual_seed(0)
    a = ns.U2.UNet(3, 4).state_dict()
    torch.manual_seed(0)
    b = B.UNet(3, 4).state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
    torch.manual_seed(0)
    a = ns.M3.UNet3D(1, 1, f_maps=16, layer_order="cbr", upsample="deconv").state_dict()
    torch.manual_seed(0)
    b = B.UNet3D(1, 1, f_maps=16, layer_order="cbr", upsample="deconv").state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    print("seeded init identical: reference vs product modules")


def gen_unet3plus(ns):
    """UNet 3+ (model/unet2d/unet.py:136-446), the default model of train.py: 27 M parameters, so the
    fixture stores the seed of oracle.unet3plus.fill_state_dict_ (a deterministic fill by state_dict key)
    instead of the weights, plus input, target, logits, loss, the gradients of a few small parameters and
    the gradient norm of every parameter."""
    from oracle import unet3plus as O3
    torch.manual_seed(0)
    net = ns.U2.UNet_3Plus(3, 1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    O3.fill_state_dict_(sd, 7)
    net.load_state_dict(sd)
    net.train()
    g = torch.Generator().manual_seed(21)
    x = torch.randn((2, 3, 32, 48), generator=g)
    y = blob_target((2, 1, 32, 48), 22)
    logits = net(x)
    loss = torch.nn.BCEWithLogitsLoss()(logits, y)
    loss.backward()
    names = [k for k, _ in net.named_parameters()]
    keep = ["outconv1.weight", "outconv1.bias", "bn1d_1.weight", "bn1d_1.bias", "conv1.conv1.0.bias",
            "conv1.conv1.1.weight", "h1_PT_hd4_bn.weight", "hd5_UT_hd1_bn.bias", "h3_Cat_hd3_conv.bias",
            "conv5.conv2.1.weight"]
    out = dict(seed=7, x=_np(x), y=_np(y), logits=_np(logits), loss=_np(loss),
               grad_names=np.array(names), grad_norms=np.array([p.grad.norm().item() for _, p in net.named_parameters()]),
               running_mean_bn1d_1=_np(net.state_dict()["bn1d_1.running_mean"]),
               running_var_conv1=_np(net.state_dict()["conv1.conv1.1.running_var"]),
               state_dict_keys=np.array(list(net.state_dict().keys())))
    grads = dict(net.named_parameters())
    for k in keep:
        out["grad::" + k] = _np(grads[k].grad)
    np.savez_compressed(os.path.join(OUT, "unet3plus_3_1.npz"), **out)
    print("unet3plus_3_1", float(loss), logits.abs().max().item())


def gen_unet3plus_deepsup(ns):
    """UNet_3Plus_DeepSup (unet.py:454-787) through the reference's own UNetModel-style loss sum
    (unet.py:1203-1208) with BCE as the criterion: the five per-scale logits, the summed loss, gradient
    norms of every parameter and the gradients of the five output convs' biases + two small tensors."""
    from oracle import unet3plus as O3
    torch.manual_seed(0)
    net = ns.U2.UNet_3Plus_DeepSup(3, 1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    O3.fill_state_dict_(sd, 11)
    net.load_state_dict(sd)
    net.train()
    g = torch.Generator().manual_seed(31)
    x = torch.randn((2, 3, 32, 32), generator=g)
    y = blob_target((2, 1, 32, 32), 32)
    ds = net(x)
    bce = torch.nn.BCEWithLogitsLoss()
    loss = sum(bce(d, y) for d in ds)
    loss.backward()
    names = [k for k, _ in net.named_parameters()]
    grads = dict(net.named_parameters())
    out = dict(seed=11, x=_np(x), y=_np(y), loss=_np(loss), grad_names=np.array(names),
               grad_norms=np.array([p.grad.norm().item() for _, p in net.named_parameters()]),
               state_dict_keys=np.array(list(net.state_dict().keys())))
    for i, d in enumerate(ds):
        out[f"d{i + 1}"] = _np(d)
    for k in [f"outconv{i}.bias" for i in range(1, 6)] + ["outconv5.weight", "bn4d_1.weight"]:
        out["grad::" + k] = _np(grads[k].grad)
    np.savez_compressed(os.path.join(OUT, "unet3plus_deepsup_3_1.npz"), **out)
    print("unet3plus_deepsup_3_1", float(loss), [float(d.abs().max()) for d in ds])


def gen_unet3plus_cgm(ns):
    """UNet_3Plus_DeepSup_CGM (unet.py:795-1153) with the dropout of its classification branch switched off
    (p = 0: a random mask cannot be pinned), training-mode BatchNorm, 4 samples so that both gate values occur;
    BCE on the five probability maps.  Stores the class probabilities so that the test can see the margin."""
    from oracle import unet3plus as O3
    torch.manual_seed(0)
    net = ns.U2.UNet_3Plus_DeepSup_CGM(3, 1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    O3.fill_state_dict_(sd, 27)      # a seed whose gates are mixed with a margin (0.087)
    net.load_state_dict(sd)
    net.train()
    net.cls[0].p = 0.0
    g = torch.Generator().manual_seed(41)
    x = torch.randn((4, 3, 32, 32), generator=g)
    y = blob_target((4, 1, 32, 32), 42)
    feats = {}
    hook = net.conv5.register_forward_hook(lambda m, i, o: feats.__setitem__("hd5", o.detach()))
    ds = net(x)
    hook.remove()
    cls = net.cls(feats["hd5"]).squeeze(3).squeeze(2)
    loss = sum(torch.nn.functional.binary_cross_entropy(d, y) for d in ds)
    loss.backward()
    names = [k for k, p in net.named_parameters()]
    has_grad = np.array([p.grad is not None for _, p in net.named_parameters()])
    out = dict(seed=27, x=_np(x), y=_np(y), loss=_np(loss), cls=_np(cls), grad_names=np.array(names), has_grad=has_grad,
               grad_norms=np.array([0.0 if p.grad is None else p.grad.norm().item() for _, p in net.named_parameters()]),
               state_dict_keys=np.array(list(net.state_dict().keys())))
    for i, d in enumerate(ds):
        out[f"d{i + 1}"] = _np(d)
    np.savez_compressed(os.path.join(OUT, "unet3plus_cgm_3_1.npz"), **out)
    print("unet3plus_cgm_3_1", float(loss), "cls", cls.detach().numpy().round(4).tolist())


def gen_predictor(ns):
    """Patch logic of the sliding-window predictor from the reference's own dataset/unet3d_dataset/utils.py
    (loaded by path: the package __init__ needs albumentations): slice lists for several volume / patch / stride
    combinations, mirror_pad and remove_padding outputs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_ref_ds_utils", os.path.join(ref_loader.REF_ROOT, "dataset", "unet3d_dataset", "utils.py"))
    U = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(U)
    cases = [((40, 70, 66), (16, 64, 64), (8, 32, 32)), ((32, 64, 64), (32, 64, 64), (32, 64, 64)),
             ((33, 65, 130), (16, 64, 64), (16, 64, 64)), ((20, 96, 100), (8, 64, 64), (5, 17, 23)),
             ((17, 64, 64), (16, 64, 64), (3, 64, 64))]
    out = {"cases": np.array(cases)}
    for i, (shape, patch, stride) in enumerate(cases):
        sl = U.SliceBuilder._build_slices(np.zeros(shape, dtype=np.uint8), patch, stride)
        out[f"slices{i}"] = np.array([[s.start for s in idx] + [s.stop for s in idx] for idx in sl])
    rng = np.random.default_rng(0)
    a3 = rng.standard_normal((6, 7, 9)).astype(np.float32)
    a4 = rng.standard_normal((2, 6, 7, 9)).astype(np.float32)
    out["pad_in3"], out["pad_in4"] = a3, a4
    out["pad_out3"] = U.mirror_pad(a3, (2, 3, 4))
    out["pad_out4"] = U.mirror_pad(a4, (1, 0, 5))
    out["unpad_out4"] = U.remove_padding(U.mirror_pad(a4, (1, 0, 5))[None], (1, 0, 5))[0]
    np.savez_compressed(os.path.join(OUT, "predictor_slices.npz"), **out)
    print("predictor_slices", [len(out[f"slices{i}"]) for i in range(len(cases))])


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load()
    check_seed_init(ns)
    gen_unet2d(ns)
    gen_unet3d(ns)
    gen_resunet3d(ns)
    gen_blocks(ns)
    gen_losses_metrics(ns)
    gen_unet3plus(ns)
    gen_unet3plus_deepsup(ns)
    gen_unet3plus_cgm(ns)
    gen_predictor(ns)


if __name__ == "__main__":
    main()
