"""Round-2 GPU parity (all through the C ABI):

* the new building blocks against the reference's own outputs (tests/golden/blocks_r2.npz, resunet_r2.npz);
* parity AT BASELINE SCALE against the fp32 CPU oracle: UNet3D(1,1) f_maps=64 at 128^3 (gcr + nearest and
  cbr + deconv, weights trained by the oracle on the blob task) and UNet(3,4) at 512^2;
* every headline layer shape against an fp32 ``F.conv*`` reference under every schedule knob of the tcgen05
  kernels (halo R = 1/2/4, weight-resident on/off, wide / narrow weight gradient, 4 / 8 epilogue warps, per-tap);
* CrossEntropy ignore_index / bad labels, gradient accumulation under the data-parallel wrapper;
* LazyPredictor and the serving handler.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import predictor as OP
from oracle import unet as O
from test_oracle_golden_r2 import case_of, cases, oracle_block  # noqa: F401

pytestmark = pytest.mark.gpu

BLOCKS = load_golden("blocks_r2")
NETS = load_golden("resunet_r2")


@pytest.fixture(scope="module")
def B():
    import mdeical_image_segmentation_b200 as B
    return B


def _make_block(B, case):
    kind, _, tail = case.partition("_")
    is3d = "2d" not in kind
    if kind.startswith("sc"):
        order = tail.replace("eval_", "")
        m = B.SingleConv(4, 8, order=order, num_groups=4, dropout_prob=0.3, is3d=is3d)
        return m.eval() if "eval" in tail else m.train()
    if kind.startswith("rbse"):
        return B.ResNetBlockSE(4, 8, order="gcr", num_groups=4, se_module=tail, is3d=True).train()
    if kind.startswith("rb"):
        order = tail.split("_")[0]
        cout = 4 if tail.endswith("same") else 8
        return B.ResNetBlock(4, cout, order=order, num_groups=4, is3d=is3d).train()
    if kind == "se":
        return {"cse": lambda: B.ChannelSELayer3D(8, 2), "sse": lambda: B.SpatialSELayer3D(8),
                "scse": lambda: B.ChannelSpatialSELayer3D(8, 2)}[tail]().train()
    if kind.startswith("dec"):
        kw = {"dec3d_nearest": dict(in_channels=24, out_channels=8, conv_layer_order="gcr", num_groups=4),
              "dec3d_nearest_ragged": dict(in_channels=24, out_channels=8, conv_layer_order="gcr", num_groups=4),
              "dec3d_deconv": dict(in_channels=16, out_channels=8, conv_layer_order="cbr", upsample="deconv"),
              "dec3d_res": dict(in_channels=16, out_channels=8, basic_module=B.ResNetBlock, conv_layer_order="gcr",
                                num_groups=4),
              "dec2d_nearest": dict(in_channels=24, out_channels=8, conv_layer_order="cr", is3d=False)}[case]
        return B.Decoder(**kw).train()
    if kind.startswith("enc"):
        kw = {"enc3d_avg": dict(in_channels=4, out_channels=8, pool_type="avg", conv_layer_order="gcr", num_groups=4),
              "enc3d_max_res": dict(in_channels=4, out_channels=8, basic_module=B.ResNetBlock,
                                    conv_layer_order="gcr", num_groups=4),
              "enc2d_avg": dict(in_channels=4, out_channels=8, pool_type="avg", conv_layer_order="cr", is3d=False)}[case]
        return B.Encoder(**kw).train()
    raise KeyError(case)


@pytest.mark.parametrize("case", cases(BLOCKS))
def test_new_blocks_vs_reference_golden(B, case):
    """Stand-alone forward + backward of every block the verdict listed as missing, against the reference."""
    sd, ins, dins, grads = case_of(BLOCKS, case)
    mod = _make_block(B, case)
    mod.load_state_dict(sd)
    mod = mod.cuda()
    names = ["skip", "x"] if case.startswith("dec") else ["x"]
    args = [ins[n].cuda().requires_grad_(True) for n in names]
    y = mod(*args)
    ref_y = torch.from_numpy(BLOCKS[case + "/y"])
    assert y.shape == ref_y.shape
    assert rel_l2(y.detach().cpu(), ref_y) < 1.2e-2, case
    (y * torch.from_numpy(BLOCKS[case + "/r"]).cuda()).sum().backward()
    # random weights on noise inputs: a fraction f of ReLU / pooling decisions sits within bf16 rounding of a tie
    # and flips, which alone moves a gradient by ~sqrt(f) — the strict 2e-2 bar is held on the trained
    # fixtures below; here the bound only has to catch a wrong formula (errors of order 1)
    # a max-pool on the fp32 noise input: storing it in bf16 makes ~1 % of the 2x2x2 windows change their
    # arg-max, which moves that window's whole gradient to another voxel
    din_tol = 0.2 if "max" in case else 0.1
    for n, a in zip(names, args):
        assert rel_l2(a.grad.cpu(), dins[n]) < din_tol, (case, n, rel_l2(a.grad.cpu(), dins[n]))
    named = dict(mod.named_parameters())
    gnorm = float(np.sqrt(sum(float((g.double() ** 2).sum()) for g in grads.values())))
    for k, g in grads.items():
        got = named[k].grad
        got = torch.zeros_like(g) if got is None else got.cpu()
        err = (got - g).norm().item()
        # conv biases in front of a norm / zero-gradient tensors: absolute floor tied to the global norm
        assert err < 1e-1 * g.norm().item() + 5e-3 * gnorm, (case, k, err, g.norm().item(), gnorm)


@pytest.mark.parametrize("case,cls,kw", [
    ("resunet2d_gcr", "ResidualUNet2D", dict(in_channels=2, out_channels=2, f_maps=[8, 16, 32], num_groups=4)),
    ("resunet3d_cge", "ResidualUNet3D", dict(in_channels=1, out_channels=1, f_maps=[8, 16, 32], num_groups=4,
                                             layer_order="cge")),
    ("resunet3d_gcl", "ResidualUNet3D", dict(in_channels=1, out_channels=2, f_maps=[8, 16], num_groups=4,
                                             layer_order="gcl", num_levels=2))])
def test_residual_networks_r2_vs_reference_golden(B, case, cls, kw):
    """ResidualUNet2D (model.py:323-359) and ELU / LeakyReLU residual networks through the whole-network executor."""
    sd = {k[len(case) + 4:]: torch.from_numpy(v.copy()) for k, v in NETS.items() if k.startswith(case + "/sd/")}
    net = getattr(B, cls)(**kw)
    net.load_state_dict(sd)
    net = net.cuda().train()
    x, t = torch.from_numpy(NETS[case + "/x"]).cuda(), torch.from_numpy(NETS[case + "/t"]).cuda()
    logits = net(x)
    loss = B.BCEDiceLoss(1.0, 1.0)(logits, t)
    loss.backward()
    assert rel_l2(logits.detach().cpu(), NETS[case + "/logits"]) < 2e-2
    assert abs(loss.item() - float(NETS[case + "/loss"])) < 2e-3
    keys = sorted(k[len(case) + 6:] for k in NETS if k.startswith(case + "/grad/"))
    named = dict(net.named_parameters())
    ga = torch.cat([named[k].grad.flatten().cpu().double() for k in keys])
    gb = torch.cat([torch.from_numpy(NETS[f"{case}/grad/{k}"]).flatten().double() for k in keys])
    assert ((ga - gb).norm() / gb.norm()).item() < 0.2     # random init + noise input (see the trained fixtures)
    net.eval()
    with torch.no_grad():
        ev = net(x)
    assert rel_l2(ev.cpu(), NETS[case + "/logits_eval"]) < 2e-2


def test_elu_residual_network_trained_parity(B):
    """The 2e-2 bar on a trained fixture for the ELU residual network (ResNetBlock's own default order cge)."""
    torch.manual_seed(0)
    kw = dict(num_levels=3, layer_order="cge", num_groups=4)
    net = B.ResidualUNet3D(1, 1, f_maps=[8, 16, 32], num_groups=4, layer_order="cge")
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    params = [v for v in sd.values() if v.is_floating_point()]
    for p in params:
        p.requires_grad_(True)
    opt = torch.optim.Adam(params, lr=1e-3)
    from test_gpu_parity import _blob_task
    for s in range(40):
        x, m = _blob_task((2, 1, 16, 16, 16), 100 + s)
        opt.zero_grad()
        O.bce_dice_loss(O.residual_unet3d_forward(sd, x, **kw), m).backward()
        opt.step()
    x, m = _blob_task((2, 1, 32, 32, 32), 999)
    for p in params:
        p.grad = None
    ref = O.residual_unet3d_forward(sd, x, **kw)
    ref_loss = O.bce_dice_loss(ref, m)
    ref_loss.backward()
    net.load_state_dict({k: v.detach() for k, v in sd.items()})
    net = net.cuda().train()
    logits = net(x.cuda())
    loss = B.BCEDiceLoss(1.0, 1.0)(logits, m.cuda())
    loss.backward()
    lg = logits.detach().cpu()
    assert rel_l2(lg, ref.detach()) < 2e-2
    assert ((lg > 0) == (ref.detach() > 0)).float().mean().item() >= 0.999
    assert abs(loss.item() - ref_loss.item()) < 1e-3
    named = dict(net.named_parameters())
    ga = torch.cat([named[k].grad.flatten().cpu().double() for k in sorted(named)])
    gb = torch.cat([sd[k].grad.flatten().double() for k in sorted(named)])
    assert ((ga - gb).norm() / gb.norm()).item() < 2e-2


def test_dropout_letters_in_training_mode(B):
    """'d' / 'D' with a live mask: outputs are the eval output times keep / (1 - p), whole channels for 'D', and
    the backward uses the same mask."""
    torch.manual_seed(5)
    x = torch.randn((2, 8, 4, 8, 8), device="cuda")
    for order, chanwise in (("crd", False), ("crD", True)):
        m = B.SingleConv(8, 16, order=order, dropout_prob=0.4).cuda()
        m.eval()
        with torch.no_grad():
            base = m(x)
        m.train()
        xin = x.clone().requires_grad_(True)
        y = m(xin)
        keep = (y != 0) | (base == 0)
        scaled = torch.where(y != 0, y / base.clamp_min(1e-20), torch.zeros_like(y))
        nz = (y != 0) & (base.abs() > 1e-3)
        assert torch.allclose(scaled[nz], torch.full_like(scaled[nz], 1 / 0.6), rtol=2e-2)
        frac = (y != 0).float().sum() / (base != 0).float().sum()
        lo, hi = (0.3, 0.9) if chanwise else (0.5, 0.7)          # 32 (n, c) coins vs 32 k element coins
        assert lo < frac.item() < hi, frac.item()
        if chanwise:
            per = (y != 0).flatten(2).any(-1) | ~(base != 0).flatten(2).any(-1)      # (n, c): kept or empty
            dropped = ~per
            assert dropped.any() and (y.flatten(2)[dropped] == 0).all()
        y.sum().backward()
        assert torch.isfinite(xin.grad).all() and xin.grad.abs().sum() > 0
        del keep


def test_cross_entropy_ignore_index_and_bad_labels(B):
    """nn.CrossEntropyLoss() ignores -100 labels (mean over the valid positions); an out-of-range
    label must never be dereferenced (NaN loss, zero gradient there); (N,1,H,W) targets are refused."""
    g = torch.Generator().manual_seed(3)
    logits = torch.randn((2, 4, 16, 16), generator=g)
    y = torch.randint(0, 4, (2, 16, 16), generator=g)
    y[0, :5] = -100
    ref_in = logits.clone().requires_grad_(True)
    ref = F.cross_entropy(ref_in, y)
    ref.backward()
    xin = logits.cuda().requires_grad_(True)
    loss = B.CrossEntropyLoss()(xin, y.cuda())
    loss.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert rel_l2(xin.grad.cpu(), ref_in.grad) < 1e-5
    assert (xin.grad[0, :, :5] == 0).all()
    bad = y.clone()
    bad[1, 3, 3] = 255
    xin2 = logits.cuda().requires_grad_(True)
    l2 = B.CrossEntropyLoss()(xin2, bad.cuda())
    assert torch.isnan(l2)
    torch.cuda.synchronize()                    # no illegal address
    with pytest.raises(ValueError):
        B.CrossEntropyLoss()(logits.cuda(), y.unsqueeze(1).cuda())
    empty = torch.full_like(y, -100)
    assert torch.isnan(B.CrossEntropyLoss()(logits.cuda(), empty.cuda()))      # 0 / 0, as the stock op


def test_gradient_accumulation_under_data_parallel_wrapper(B):
    """two backwards without zero_grad under DataParallelUNet accumulate old + new (the reducer must
    not hand out views of its persistent bucket), and zero_grad(set_to_none=False) keeps gradients correct."""
    from mdeical_image_segmentation_b200.parallel import DataParallelUNet
    torch.manual_seed(1)
    net = B.UNet3D(1, 1, f_maps=[8, 16], num_levels=2, num_groups=4).cuda().train()
    dp = DataParallelUNet(net, bucket_mb=0.01)          # several buckets, flushes land on norm parameters too
    crit = B.BCEDiceLoss(1.0, 1.0)
    g = torch.Generator().manual_seed(2)
    xs = [torch.randn((1, 1, 16, 16, 16), generator=g).cuda() for _ in range(2)]
    ts = [(torch.rand((1, 1, 16, 16, 16), generator=g) > 0.6).float().cuda() for _ in range(2)]
    single = []
    for x, t in zip(xs, ts):
        net.zero_grad(set_to_none=True)
        crit(dp(x), t).backward()
        single.append({k: p.grad.clone() for k, p in net.named_parameters()})
    net.zero_grad(set_to_none=True)
    for x, t in zip(xs, ts):
        crit(dp(x), t).backward()
    for k, p in net.named_parameters():
        want = single[0][k] + single[1][k]
        assert (p.grad - want).norm() <= 2e-2 * want.norm() + 1e-5, k
    net.zero_grad(set_to_none=False)
    crit(dp(xs[0]), ts[0]).backward()
    for k, p in net.named_parameters():
        assert (p.grad - single[0][k]).norm() <= 2e-2 * single[0][k].norm() + 1e-5, k


def test_hf_wrapper_3d_gradients_vs_oracle(B):
    """the double-sigmoid HF wrapper (UNet3D.py:134-154) — gradients against the oracle."""
    torch.manual_seed(9)
    cfg = B.UNet3DForMedicalSegmentationConfig(in_channels=1, out_channels=1, f_maps=[8, 16], num_groups=4,
                                               num_levels=2)
    model = B.UNet3DForMedicalSegmentation(cfg)
    sd = {k[len("model."):]: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(10)
    x = torch.randn((2, 1, 16, 16, 16), generator=g)
    t = (torch.rand((2, 1, 16, 16, 16), generator=g) > 0.5).float()
    ref_loss = O.bce_dice_loss(torch.sigmoid(O.unet3d_forward(sd, x, num_levels=2, num_groups=4, training=True)), t)
    ref_loss.backward()
    model = model.cuda().train()
    out = model(volume=x.cuda(), target=t.cuda())
    out.loss.backward()
    named = {k[len("model."):]: p for k, p in model.named_parameters()}
    ga = torch.cat([named[k].grad.flatten().cpu().double() for k in sorted(named)])
    gb = torch.cat([sd[k].grad.flatten().double() for k in sorted(named)])
    assert ((ga - gb).norm() / gb.norm()).item() < 0.15          # random init + noise (bf16 bound)
    cos = (ga @ gb / (ga.norm() * gb.norm())).item()
    assert cos > 0.99


# ===================================================================================== BASELINE scale
def _oracle_threads():
    torch.set_num_threads(max(1, os.cpu_count() or 1))


@pytest.mark.parametrize("order,upsample", [("gcr", "default"), ("cbr", "deconv")])
def test_cfg3_full_width_128cubed_vs_oracle(B, order, upsample):
    """BASELINE configs[2] at its real shape: UNet3D(1,1), f_maps 64..512, 128^3, batch 1 — forward + BCEDice +
    backward against the fp32 CPU oracle.  The weights are first trained by the oracle (full width, 32^3 crops of
    the blob task), so the north star's tolerances apply: logits 2e-2 (rel-L2 and max-norm), masks >= 99.9 %,
    loss / Dice 1e-3, concatenated parameter gradient 2e-2.  This is the only place the benchmark's schedules
    (halo super-tiles R = 2/4 at 128^3, wide weight gradient, im2col first layer, 8-warp epilogues) meet the oracle."""
    from test_gpu_parity import _blob_task
    _oracle_threads()
    torch.manual_seed(0)
    net = B.UNet3D(1, 1, layer_order=order, upsample=upsample)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    kw = dict(num_levels=4, layer_order=order, num_groups=8, upsample=upsample, is3d=True)
    params = [v for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    for p in params:
        p.requires_grad_(True)
    opt = torch.optim.Adam(params, lr=1e-3)
    for s in range(24):
        x, m = _blob_task((2, 1, 32, 32, 32), 100 + s)
        opt.zero_grad()
        O.bce_dice_loss(O.unet3d_forward(sd, x, training=True, **kw), m).backward()
        opt.step()
    for p in params:
        p.grad = None
    x, m = _blob_task((1, 1, 128, 128, 128), 999)
    sd_dev = {k: v.detach().clone() for k, v in sd.items()}
    ref = O.unet3d_forward(sd, x, training=True, **kw)
    ref_loss = O.bce_dice_loss(ref, m)
    ref_loss.backward()
    ref_l = ref.detach()
    net.load_state_dict(sd_dev)
    net = net.cuda().train()
    logits = net(x.cuda())
    loss = B.BCEDiceLoss(1.0, 1.0)(logits, m.cuda())
    loss.backward()
    lg = logits.detach().cpu()
    r = rel_l2(lg, ref_l)
    mx = ((lg - ref_l).abs().max() / ref_l.abs().max()).item()
    agree = ((lg > 0) == (ref_l > 0)).float().mean().item()
    named = dict(net.named_parameters())
    keys = [k for k in sorted(named) if sd[k].grad is not None]
    ga = torch.cat([named[k].grad.flatten().cpu().double() for k in keys])
    gb = torch.cat([sd[k].grad.flatten().double() for k in keys])
    grel = ((ga - gb).norm() / gb.norm()).item()
    d_ref = O.dice_coefficient(torch.sigmoid(ref_l), m).item()
    d_dev = B.metrics.DiceCoefficient()(torch.sigmoid(logits.detach()), m.cuda()).item()
    print(f"cfg3 {order}+{upsample} 128^3: logits rel-L2 {r:.3e} max {mx:.3e} masks {agree:.5f} "
          f"loss {loss.item():.5f} / {ref_loss.item():.5f} grad rel {grel:.3e} dice {d_dev:.5f} / {d_ref:.5f}")
    assert r < 2e-2 and mx < 2e-2, (r, mx)
    assert agree >= 0.999, agree
    assert abs(loss.item() - ref_loss.item()) < 1e-3
    assert abs(d_ref - d_dev) < 1e-3
    assert grel < 2e-2, grel


def test_cfg2_unet_512_vs_oracle(B):
    """BASELINE configs[1] at its real resolution: UNet(3,4), 512^2, batch 2, CrossEntropy, against the oracle
    (weights trained by the oracle on 64^2 crops of a 4-class blob task)."""
    _oracle_threads()
    torch.manual_seed(0)
    net = B.UNet(3, 4)
    sd = {k: v.clone().requires_grad_(True) for k, v in net.state_dict().items()}
    opt = torch.optim.Adam(list(sd.values()), lr=2e-4)

    def task(shape, seed):
        g = torch.Generator().manual_seed(seed)
        blur = F.avg_pool2d(torch.randn((shape[0], 4) + shape[2:], generator=g), 9, stride=1, padding=4)
        y = blur.argmax(1)
        x = 0.5 * torch.randn(shape, generator=g) + F.one_hot(y, 4).movedim(-1, 1)[:, :3].float()
        return x, y

    for s in range(12):
        x, y = task((2, 3, 64, 64), 100 + s)
        opt.zero_grad()
        F.cross_entropy(O.unet2d_forward(sd, x), y).backward()
        opt.step()
    for v in sd.values():
        v.grad = None
    x, y = task((2, 3, 512, 512), 999)
    ref = O.unet2d_forward(sd, x)
    ref_loss = F.cross_entropy(ref, y)
    ref_loss.backward()
    net.load_state_dict({k: v.detach() for k, v in sd.items()})
    net = net.cuda().train()
    logits = net(x.cuda())
    loss = B.CrossEntropyLoss()(logits, y.cuda())
    loss.backward()
    lg, ref_l = logits.detach().cpu(), ref.detach()
    r = rel_l2(lg, ref_l)
    agree = (lg.argmax(1) == ref_l.argmax(1)).float().mean().item()
    named = dict(net.named_parameters())
    ga = torch.cat([named[k].grad.flatten().cpu().double() for k in sorted(named)])
    gb = torch.cat([sd[k].grad.flatten().double() for k in sorted(named)])
    grel = ((ga - gb).norm() / gb.norm()).item()
    print(f"cfg2 512^2: logits rel-L2 {r:.3e} argmax agreement {agree:.5f} loss {loss.item():.5f} / "
          f"{ref_loss.item():.5f} grad rel {grel:.3e}")
    assert r < 2e-2, r
    assert agree >= 0.999, agree
    assert abs(loss.item() - ref_loss.item()) < 1e-3
    assert grel < 2e-2, grel


# ===================================================================================== headline layers x knobs
KNOBS = {                      # b200_set_tuning(key, value) lists; () = defaults
    "default": (),
    "halo_R1": ((2, 1),),
    "halo_R2": ((2, 2),),
    "streamed_weights": ((3, 0),),
    "epi4": ((5, 4),),
    "epi8": ((5, 8),),
    "narrow_wgrad": ((6, 0), (8, 0)),
    "pair_wgrad": ((8, 0),),           # the tap-pair weight gradient instead of the 3-plane N = 192 kernel
    "per_tap": ((0, 0), (1, 0)),
    "roll": ((10, 2),),                # the experimental depth-rolling N = 3 C_out kernel (C_out <= 64 layers)
    "roll_seg5": ((10, 2), (11, 5)),   # ... with ragged depth segments (128 = 25 x 5 + 3)
}
DEFAULTS = {0: 1, 1: 1, 2: 4, 3: 1, 4: 0, 5: 0, 6: 1, 7: 0, 8: 1, 9: 80, 10: 0, 11: 0}

LAYERS = [          # name, n, d, h, w, c_in, c_out  (SURVEY App. A headline shapes; batch cut to bound memory)
    ("3d_192_64_128", 1, 128, 128, 128, 192, 64),
    ("3d_64_64_128", 1, 128, 128, 128, 64, 64),
    ("3d_384_128_64", 1, 64, 64, 64, 384, 128),
    ("3d_768_256_32", 2, 32, 32, 32, 768, 256),
    ("2d_1024_1024_32", 8, 1, 32, 32, 1024, 1024),
    ("2d_128_64_512", 2, 1, 512, 512, 128, 64),
]


def _rel_dev(a, b):
    """rel-L2 on the device (the headline tensors are GBs: no host round trip)."""
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _set_knobs(lib, pairs):
    for k, v in DEFAULTS.items():
        lib.b200_set_tuning(k, v)
    for k, v in pairs:
        lib.b200_set_tuning(k, v)


@pytest.mark.parametrize("layer", LAYERS, ids=[l[0] for l in LAYERS])
def test_headline_layers_under_every_schedule_vs_fp32_reference(B, layer):
    """forward (bias + ReLU), data gradient (with ReLU mask) and weight gradient (with fused bias gradient) of each
    headline layer, under every schedule knob, against fp32 ``F.conv*`` on the same bf16-rounded operands."""
    from mdeical_image_segmentation_b200 import _lib, ops
    from mdeical_image_segmentation_b200.ops import CL
    name, n, d, h, w, cin, cout = layer
    is3d = d > 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(11)
    x = CL((torch.randn((n, d, h, w, cin), generator=g)).to(torch.bfloat16).cuda())
    dy = CL((torch.randn((n, d, h, w, cout), generator=g)).to(torch.bfloat16).cuda())
    taps = 27 if is3d else 9
    wshape = (cout, cin, 3, 3, 3) if is3d else (cout, cin, 3, 3)
    wt = (torch.randn(wshape, generator=g) / (cin * taps) ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    wq = wt.to(torch.bfloat16).float()
    conv = F.conv3d if is3d else F.conv2d
    sq = (lambda t: t) if is3d else (lambda t: t[:, :, 0])
    with torch.no_grad():
        xin = sq(x.to_nchw())
        y_ref = F.relu(conv(xin, wq, bias, padding=1))
        g_out = sq(dy.to_nchw())
        dx_ref = torch.nn.grad.conv3d_input(xin.shape, wq, g_out, padding=1) if is3d else \
            torch.nn.grad.conv2d_input(xin.shape, wq, g_out, padding=1)
        dx_ref = dx_ref * (xin > 0)
        dw_ref = torch.nn.grad.conv3d_weight(xin, wq.shape, g_out, padding=1) if is3d else \
            torch.nn.grad.conv2d_weight(xin, wq.shape, g_out, padding=1)
        db_ref = g_out.sum(dim=(0, 2, 3, 4) if is3d else (0, 2, 3))
    del xin, g_out
    lib = _lib.lib()
    wp_fwd, wp_bwd, k = ops.conv_repack_weight(wt)
    results = {}
    try:
        for knob, pairs in KNOBS.items():
            _set_knobs(lib, pairs)
            y = ops.conv_fwd(x, wp_fwd, k, bias, True, impl=1)
            dx = ops.conv_bwd_data(dy, wp_bwd, k, mask=x, impl=1)
            dw, db = ops.conv_bwd_weight(x, dy, k, tuple(wt.shape), impl=1, bias_grad=True)
            torch.cuda.synchronize()
            ey = _rel_dev(sq(y.to_nchw()), y_ref)
            ex = _rel_dev(sq(dx.to_nchw()), dx_ref)
            ew = _rel_dev(dw, dw_ref)
            eb = _rel_dev(db, db_ref)
            results[knob] = (ey, ex, ew, eb)
            del y, dx, dw, db
    finally:
        _set_knobs(lib, ())
    print(name, json.dumps({k: [f"{e:.2e}" for e in v] for k, v in results.items()}))
    for knob, (ey, ex, ew, eb) in results.items():
        # outputs are stored in bf16 (one rounding, 2^-9 relative); weight / bias gradients accumulate in fp32
        assert ey < 5e-3 and ex < 5e-3, (name, knob, ey, ex)
        assert ew < 2e-3 and eb < 2e-3, (name, knob, ew, eb)


def test_headline_transposed_convs_vs_fp32_reference(B):
    """ConvTranspose2d k2 s2 (1024 -> 512 at 32^2, layers.py:165) and ConvTranspose3d k3 s2 p1 (128 -> 64 at 64^3 ->
    127^3, buildingblocks.py:710-717): forward, data gradient and weight gradient against fp32 ``F.conv_transpose*``."""
    from mdeical_image_segmentation_b200 import ops
    from mdeical_image_segmentation_b200.ops import CL
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator().manual_seed(12)
    # ---- 2D k2 s2 with bias into a concat slice
    n, h, w, cin, cout = 4, 32, 32, 1024, 512
    x = CL(torch.randn((n, 1, h, w, cin), generator=g).to(torch.bfloat16).cuda())
    wt = (torch.randn((cin, cout, 2, 2), generator=g) / cin ** 0.5).cuda()
    bias = torch.randn(cout, generator=g).cuda()
    dy = CL(torch.randn((n, 1, 2 * h, 2 * w, cout), generator=g).to(torch.bfloat16).cuda())
    wp_fwd, wp_bwd = ops.convt_repack_weight(wt)
    y = ops.convt_fwd(x, wp_fwd, bias)
    dx = ops.convt_bwd_data(dy, wp_bwd)
    dw = ops.convt_bwd_weight(x, dy)
    wq = wt.to(torch.bfloat16).float()
    xin = x.to_nchw()[:, :, 0].requires_grad_(True)
    wq.requires_grad_(True)
    y_ref = F.conv_transpose2d(xin, wq, bias, stride=2)
    y_ref.backward(dy.to_nchw()[:, :, 0])
    assert _rel_dev(y.to_nchw()[:, :, 0], y_ref.detach()) < 5e-3
    assert _rel_dev(dx.to_nchw()[:, :, 0], xin.grad) < 5e-3
    assert _rel_dev(dw, wq.grad) < 2e-3
    # ---- 3D k3 s2 p1 (parity planes) + nearest resize to 2x
    n, s, cin, cout = 1, 64, 128, 64
    x = CL(torch.randn((n, s, s, s, cin), generator=g).to(torch.bfloat16).cuda())
    wt = (torch.randn((cin, cout, 3, 3, 3), generator=g) / (cin * 27 / 8) ** 0.5).cuda()
    wp_fwd, wp_bwd, kd = ops.deconv_repack_weight(wt)
    planes = ops.deconv_fwd(x, wp_fwd, kd)
    up = ops.deconv_resize_fwd(planes, n, kd, (2 * s, 2 * s, 2 * s))
    wq = wt.to(torch.bfloat16).float().requires_grad_(True)
    xin = x.to_nchw().requires_grad_(True)
    ref = F.interpolate(F.conv_transpose3d(xin, wq, None, stride=2, padding=1), size=(2 * s,) * 3)
    assert _rel_dev(up.to_nchw(), ref.detach()) < 5e-3
    dup = CL(torch.randn((n, 2 * s, 2 * s, 2 * s, cout), generator=g).to(torch.bfloat16).cuda())
    ref.backward(dup.to_nchw())
    d_planes = ops.deconv_resize_bwd(dup, x.dims, kd)
    dw = ops.deconv_bwd_weight(x, d_planes, tuple(wt.shape), kd)
    dx = ops.deconv_bwd_data(d_planes, wp_bwd, x.dims, kd)
    # d_planes is rounded to bf16 after summing up to 8 resized children: one extra rounding on this path
    assert _rel_dev(dx.to_nchw(), xin.grad) < 8e-3
    assert _rel_dev(dw, wq.grad) < 5e-3


# ===================================================================================== predictor / serving
def test_lazy_predictor_vs_oracle(B, tmp_path):
    """LazyPredictor (predictor.py:171-223): bit-identical to the numpy restatement of the reference loop, with
    halos, ragged last patches, channel selection and block-wise segmentation; store = .npy memory maps."""
    from mdeical_image_segmentation_b200.predictor import LazyPredictor

    class Stub(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.p = torch.nn.Parameter(torch.zeros(1))

        def forward(self, b):
            pos = torch.arange(b.shape[-1], device=b.device, dtype=torch.float32)
            return torch.stack([torch.tanh(b.sum(1)) + 0.01 * pos, torch.cos(b[:, 0]) - 0.02 * pos, 0.5 * b[:, -1]], 1)

    def model_fn(batch):
        with torch.no_grad():
            return Stub()(torch.from_numpy(np.asarray(batch, dtype=np.float32))).numpy()

    rng = np.random.default_rng(1)
    raw = rng.standard_normal((2, 11, 13, 10)).astype(np.float32)
    model = Stub().cuda()
    for seg, chan, out_dir in ((False, None, None), (True, None, str(tmp_path)), (False, 1, str(tmp_path))):
        lp = LazyPredictor(model, out_dir, out_channels=3, save_segmentation=seg, prediction_channel=chan)
        got = lp.predict_volume(raw, (6, 8, 6), (3, 5, 4), (1, 2, 1), batch_size=2)
        want = OP.lazy_predict(model_fn, raw, (6, 8, 6), (3, 5, 4), (1, 2, 1), out_channels=3,
                               prediction_channel=chan, save_segmentation=seg, batch_size=2)
        assert got.shape == want.shape
        if seg:
            assert np.array_equal(got, want)
        else:
            assert np.allclose(got, want, rtol=0, atol=2e-6)     # device tanh / cos vs numpy's
    # same numbers as the in-memory predictor
    sp = B.StandardPredictor(model, None, out_channels=3)
    a = sp.predict_volume(raw, (6, 8, 6), (3, 5, 4), (1, 2, 1), batch_size=2)
    b = LazyPredictor(model, None, out_channels=3).predict_volume(raw, (6, 8, 6), (3, 5, 4), (1, 2, 1), batch_size=2)
    assert np.array_equal(a, b)


def test_serving_handler_contract(B, tmp_path):
    """api/api.py:52-82 for a segmentation model: POST {"image_path"} -> preprocess -> model -> mask PNG written,
    its path returned with status 200; batched requests share one bucketed CUDA-graph replay; the mask equals the
    oracle's on the same preprocessed input."""
    from PIL import Image
    from mdeical_image_segmentation_b200.serving import SegmentationService, create_app
    torch.manual_seed(3)
    net = B.UNet(3, 2)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda().eval()
    rng = np.random.default_rng(0)
    paths = []
    for i in range(3):
        img = (rng.random((300 + 10 * i, 280, 3)) * 255).astype(np.uint8)
        img[100:200, 80:200] = 255 - img[100:200, 80:200] // 4
        p = tmp_path / f"img{i}.png"
        Image.fromarray(img).save(p)
        paths.append(str(p))
    svc = SegmentationService(net, str(tmp_path / "masks"), in_channels=3, max_batch=4)
    ans = svc.handle({"image_path": paths[0]})
    assert ans["status"] == 200 and os.path.exists(ans["response"]) and "time" in ans
    mask = np.asarray(Image.open(ans["response"]))
    assert mask.shape == (224, 224)
    x = svc.build_input(paths[0])
    assert tuple(x.shape) == (3, 224, 224)
    with torch.no_grad():
        ref = O.unet2d_forward(sd, x[None]).argmax(1)[0].numpy()
    got = (mask > 0).astype(np.int64) if mask.max() > 1 else mask.astype(np.int64)
    assert (got == ref).mean() >= 0.995          # random-init logits are close to ties on flat regions
    many = svc.handle_many([{"image_path": p} for p in paths] + [{"nope": 1}])
    assert [a["status"] for a in many] == [200, 200, 200, 400]
    again = np.asarray(Image.open(many[0]["response"]))
    assert np.array_equal(again, mask)           # batch-3 (bucket 4) replay gives the batch-1 mask
    assert set(k[0] for k in svc._graphs) == {1, 4}
    pct = svc.latency_percentiles()
    assert pct["n"] == 2 and pct["p50_ms"] > 0
    app = create_app(svc)
    assert any(getattr(r, "path", None) == "/" for r in app.routes)


# ===================================================================================== augmentation
def test_gpu_augmentations_vs_reference_golden(B):
    """The device transforms with the reference's class names against outputs of the reference's own classes for the
    same RandomState seeds: flips / rotations bit-exact, the arithmetic ones to fp32 rounding, the elastic deformation
    (host-noise mode = the reference's random stream; device Gaussian smoothing + tri-linear / nearest gather) to 1e-4."""
    from mdeical_image_segmentation_b200 import augment as A
    aug = load_golden("augment_r2")
    m4 = torch.from_numpy(aug["m4"]).cuda()
    m3 = torch.from_numpy(aug["m3"]).cuda()
    lab = torch.from_numpy(aug["lab"]).cuda()
    for seed in range(6):
        assert np.array_equal(A.RandomFlip(np.random.RandomState(seed))(m4).cpu().numpy(), aug[f"flip/{seed}"])
        assert np.array_equal(A.RandomRotate90(np.random.RandomState(seed))(m4).cpu().numpy(), aug[f"rot/{seed}"])
        assert np.array_equal(A.RandomRotate90(np.random.RandomState(seed))(m3).cpu().numpy(), aug[f"rot3/{seed}"])
        got = A.RandomContrast(np.random.RandomState(seed), execution_probability=0.7)(m4).cpu().numpy()
        assert np.allclose(got, aug[f"contrast/{seed}"], atol=1e-6)
    src = {"m4": m4, "m3": m3, "lab": lab}
    for key in [k for k in aug if k.startswith("elastic/")]:
        seed, order, name = key.split("/")[1].split("_")
        e = A.ElasticDeformation(np.random.RandomState(int(seed)), spline_order=int(order), alpha=15, sigma=2,
                                 execution_probability=1.0, noise="host")
        got = e(src[name]).cpu().numpy()
        if int(order) == 0:      # nearest: identical except where a coordinate sits on a rounding boundary
            assert (got == aug[key]).mean() > 0.999, key
        else:
            assert np.abs(got - aug[key]).max() < 1e-4, (key, np.abs(got - aug[key]).max())
    e2 = A.ElasticDeformation(np.random.RandomState(4), spline_order=1, alpha=15, sigma=2, execution_probability=1.0,
                              apply_3d=False, noise="host")
    assert np.abs(e2(m3).cpu().numpy() - aug["elastic2d/4_1_m3"]).max() < 1e-4
    assert np.allclose(A.Standardize()(m4).cpu().numpy(), aug["standardize/global"], atol=2e-5)
    assert np.allclose(A.Standardize(channelwise=True)(m4).cpu().numpy(), aug["standardize/channelwise"], atol=2e-5)
    assert np.allclose(A.Standardize(mean=0.1, std=0.5)(m4).cpu().numpy(), aug["standardize/fixed"], atol=2e-6)
    assert np.allclose(A.Normalize()(m4 * 3).cpu().numpy(), aug["normalize/default"], atol=2e-6)
    assert np.allclose(A.Normalize(norm01=True)(m4 * 3).cpu().numpy(), aug["normalize/norm01"], atol=2e-6)
    assert np.allclose(A.Normalize(min_value=-2.0, max_value=2.5)(m4 * 3).cpu().numpy(), aug["normalize/fixed"], atol=2e-6)
    pipe = A.Compose([A.Standardize(), A.RandomFlip(np.random.RandomState(5)), A.RandomRotate90(np.random.RandomState(5)),
                      A.RandomContrast(np.random.RandomState(5), execution_probability=1.0),
                      A.ElasticDeformation(np.random.RandomState(5), spline_order=1, alpha=15, sigma=2,
                                           execution_probability=1.0, noise="host")])
    out = pipe(m4).cpu().numpy()
    assert out.shape == aug["pipeline"].shape and np.abs(out - aug["pipeline"]).max() < 2e-4
    # device-noise mode and the reference's default strength on a patch-sized volume: finite, same shape, deformed
    big = torch.rand((1, 32, 96, 96), device="cuda")
    ed = A.ElasticDeformation(np.random.RandomState(0), spline_order=1, alpha=2000, sigma=50, execution_probability=1.0)
    y = ed(big)
    assert y.shape == big.shape and torch.isfinite(y).all() and not torch.equal(y, big)
    n = A.AdditiveGaussianNoise(np.random.RandomState(0), execution_probability=1.0)(big)
    assert n.shape == big.shape and (n - big).std().item() > 0


# --------------------------------------------------------------- decoder concat whose up-sampled half is never written
def test_gn_apply_up_and_its_backward_match_the_materialised_concat(B):
    """b200_gn_apply_up == nearest up-sampling + b200_gn_apply on the written concat (bit-exact: same arithmetic);
    b200_gn_bwd_fused_up fed with (sum dy, sum dy * y) and the low-resolution tensor == the same backward fed with
    (sum dy, sum dy * x) and the written concat — including channels whose gamma is 0 or tiny against beta, whose
    dgamma comes from the direct reduction."""
    from mdeical_image_segmentation_b200 import ops
    g = torch.Generator().manual_seed(31)
    for dims, low_dims in (((2, 16, 24, 16), (8, 12, 8)), ((2, 1, 32, 24), (1, 16, 12)), ((1, 8, 16, 16), (8, 8, 16))):
        n, d, h, w = dims
        cs, cx, groups = 32, 64, 8
        skip = torch.randn((n, cs, d, h, w), generator=g).cuda()
        low = torch.randn((n, cx) + low_dims, generator=g).relu().cuda()
        cat = ops.CL.empty(n, d, h, w, cs + cx, skip.device)
        ops.from_nchw(skip, out=cat.slice(0, cs))
        low_cl = ops.from_nchw(low)
        full = ops.CL.empty(n, d, h, w, cs + cx, skip.device)
        ops.copy_channels(cat.slice(0, cs), full.slice(0, cs))
        ops.upsample_nearest_fwd(low_cl, (d, h, w), out=full.slice(cs, cx))
        cat.buf.view(-1, cs + cx)[:, cs:] = float("nan")      # the up-sampled half of `cat` is never written: poison it
        stats = ops.gn_stats(full, groups, 1e-5)
        gamma = (1.0 + 0.3 * torch.randn(cs + cx, generator=g)).cuda()
        beta = (0.2 * torch.randn(cs + cx, generator=g)).cuda()
        gamma[3], gamma[40], beta[40] = 0.0, 1e-3, 0.5        # degenerate channels: skip half, up-sampled half
        gamma[70], beta[70] = 0.0, 0.0
        for relu in (False, True):
            want = ops.gn_apply(full, groups, stats, gamma, beta, relu)
            got = ops.gn_apply_up(cat, cs, low_cl, groups, stats, gamma, beta, relu)
            assert torch.equal(want.dense(), got.dense()), (dims, relu)
        y = ops.gn_apply_up(cat, cs, low_cl, groups, stats, gamma, beta, False)
        # dgrad of a 3x3(x3) conv 96 <- 64 with (sum dx, sum dx * .) in its epilogue
        k = (3 if d > 1 else 1, 3, 3)
        wt = (0.05 * torch.randn((64, cs + cx) + k, generator=g)).cuda()
        _, wp_bwd, kk = ops.conv_repack_weight(wt, need_bwd=True)
        dy = ops.from_nchw(torch.randn((n, 64, d, h, w), generator=g).cuda())
        dx0, s_x = ops.conv_bwd_data(dy, wp_bwd, kk, impl=1, stats="sum", dot=full)
        dx1, s_y = ops.conv_bwd_data(dy, wp_bwd, kk, impl=1, stats="sum", dot=y)
        assert torch.equal(dx0.dense(), dx1.dense())
        for mask in (False, True):
            a = ops.gn_bwd_fused_up(full, groups, stats, gamma, dx0, s_x, cs, (n,) + low_dims, mask_relu=mask)
            b = ops.gn_bwd_fused_up(cat, groups, stats, gamma, dx1, s_y, cs, (n,) + low_dims, mask_relu=mask,
                                    x_low=low_cl, beta_if_y=beta)
            for name, u, v in zip(("dskip", "dlow", "dgamma", "dbeta"), a, b):
                u, v = (u.dense().float(), v.dense().float()) if isinstance(u, ops.CL) else (u, v)
                assert torch.isfinite(v).all(), (dims, name)
                assert rel_l2(v.cpu(), u.cpu()) < 1e-2, (dims, mask, name, rel_l2(v.cpu(), u.cpu()))
            dg_a, dg_b = a[2].cpu(), b[2].cpu()
            for ch in (3, 40, 70):       # the direct reduction is as exact as the reference path
                assert abs(dg_a[ch] - dg_b[ch]) <= 2e-3 * abs(dg_a[ch]) + 1e-3 * dg_a.abs().mean(), (dims, ch)


@pytest.mark.parametrize("dims", [(2, 1, 32, 32, 32), (1, 1, 24, 40, 16)])
def test_unet3d_with_and_without_the_unwritten_upsampled_half(B, dims):
    """UNet3D gcr + nearest: the forward that never writes the up-sampled half of the decoder concats
    (engine.FOLD_UPSAMPLE_FWD) against the one that does — same logits bit for bit, same gradients up to the
    summation order of the fused statistics — and against the fp32 oracle."""
    from mdeical_image_segmentation_b200 import engine
    torch.manual_seed(41)
    net = B.UNet3D(1, 2, f_maps=[16, 32, 64], num_groups=4, layer_order="gcr").cuda().train()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(42)
    x = torch.randn(dims, generator=g)
    t = (torch.rand((dims[0], 2) + dims[2:], generator=g) > 0.6).float()
    runs = {}
    old = engine.FOLD_UPSAMPLE_FWD
    try:
        for fold in (True, False):
            engine.FOLD_UPSAMPLE_FWD = fold
            net.zero_grad(set_to_none=True)
            out = net(x.cuda())
            loss = B.BCEDiceLoss(1.0, 1.0)(out, t.cuda())
            loss.backward()
            runs[fold] = (out.detach().clone(), torch.cat([p.grad.flatten() for p in net.parameters()]).clone())
            with torch.no_grad():
                net.eval()
                assert torch.equal(net(x.cuda()), out.detach())     # inference path folds the same way
                net.train()
    finally:
        engine.FOLD_UPSAMPLE_FWD = old
    assert torch.equal(runs[True][0], runs[False][0])
    # not bit-equal: the fused statistics are fp64 atomics of fp32 partials (order varies run to run), and a random-init
    # gradient amplifies a flipped bf16 rounding — two runs of the SAME setting differ as much (measured 1e-7 .. 1.4e-3)
    assert rel_l2(runs[True][1].cpu(), runs[False][1].cpu()) < 5e-3
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    ref = O.unet3d_forward(sd, x, training=True, num_levels=3, layer_order="gcr", num_groups=4, is3d=True)
    O.bce_dice_loss(ref, t).backward()
    gb = torch.cat([sd[k].grad.flatten() for k, _ in net.named_parameters()])
    assert rel_l2(runs[True][0].cpu(), ref.detach()) < 5e-2
    assert rel_l2(runs[True][1].cpu(), gb) < 0.3      # random init bound of test_gpu_more (strict one: trained tests)


# --------------------------------------------------------------------------- first layer on the CUDA cores (C_in = 1)
@pytest.mark.parametrize("dims", [(2, 8, 16, 32), (1, 5, 9, 37), (1, 2, 2, 2), (2, 12, 8, 70)])
def test_first_layer_direct_kernel_vs_fp32_reference(B, dims):
    """b200_conv3d_c1_fwd against F.conv3d on the same bf16-rounded operands: output within one bf16 rounding, and
    its fp64 column statistics."""
    from mdeical_image_segmentation_b200 import ops
    n, d, h, w = dims
    g = torch.Generator().manual_seed(sum(dims))
    x = torch.randn((n, 1, d, h, w), generator=g).bfloat16().float()
    wt = (0.3 * torch.randn((32, 1, 3, 3, 3), generator=g))
    bias = torch.randn(32, generator=g)
    xcl = ops.from_nchw(x.cuda(), pad=True)
    assert xcl.c == 1 and xcl.zpad == 15
    wr = wt.bfloat16().float()
    for relu in (False, True):
        ref = F.conv3d(x, wr, bias, padding=1)
        ref = ref.relu() if relu else ref
        y, cs = ops.conv3d_c1_fwd(xcl, wt.cuda(), bias.cuda(), relu, stats=True)
        got = y.to_nchw().cpu()
        assert (got - ref).abs().max() <= 2 ** -7 * ref.abs().max() + 1e-6, (dims, relu)
        want_cs = torch.stack([got.double().sum(dim=(2, 3, 4)), got.double().square().sum(dim=(2, 3, 4))], 1)
        assert torch.allclose(cs.cpu(), want_cs, rtol=1e-5, atol=1e-4)


def test_first_layer_direct_path_matches_the_tensor_core_path(B):
    """UNet3D(1, 2) f_maps 64: the network with engine.FIRST_LAYER_DIRECT on and off — logits within bf16 rounding of
    each other, first-layer parameter gradients (conv weight, its norm's weight / bias) equal within tolerance."""
    from mdeical_image_segmentation_b200 import engine
    torch.manual_seed(51)
    net = B.UNet3D(1, 2, f_maps=[64, 128], num_levels=2, num_groups=8, layer_order="gcr").cuda().train()
    g = torch.Generator().manual_seed(52)
    x = torch.randn((1, 1, 24, 32, 40), generator=g)
    t = (torch.rand((1, 2, 24, 32, 40), generator=g) > 0.6).float()
    runs = {}
    old = engine.FIRST_LAYER_DIRECT
    try:
        for direct in (True, False):
            engine.FIRST_LAYER_DIRECT = direct
            net._b200_pack_plan = None           # the set of tensor-core filter packs differs between the two
            net.zero_grad(set_to_none=True)
            out = net(x.cuda())
            B.BCEDiceLoss(1.0, 1.0)(out, t.cuda()).backward()
            runs[direct] = (out.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()})
    finally:
        engine.FIRST_LAYER_DIRECT = old
    assert rel_l2(runs[True][0].cpu(), runs[False][0].cpu()) < 5e-3
    # the two first convs differ in summation order only, but one flipped bf16 rounding of their output is amplified
    # by the random-init network behind it: whole-gradient agreement, and a looser per-tensor bound
    ga = torch.cat([v.flatten() for v in runs[True][1].values()]).cpu()
    gb = torch.cat([v.flatten() for v in runs[False][1].values()]).cpu()
    assert rel_l2(ga, gb) < 5e-2, rel_l2(ga, gb)
    for k in runs[True][1]:
        a, b = runs[True][1][k], runs[False][1][k]
        assert (a - b).norm() <= 0.15 * b.norm() + 1e-3 * gb.norm(), (k, ((a - b).norm() / b.norm()).item())


def test_maxpool_with_fused_statistics(B):
    """b200_maxpool2_fwd_sums: the pooled tensor bit-equal to b200_maxpool2_fwd, its (sum, sum of squares) per sample
    and channel equal to a direct reduction; odd extents fall back to the plain kernel (no statistics)."""
    from mdeical_image_segmentation_b200 import ops
    g = torch.Generator().manual_seed(61)
    for shape, pd in (((2, 16, 8, 12, 20), 2), ((3, 64, 1, 16, 24), 1), ((1, 8, 4, 6, 2), 2)):
        x = ops.from_nchw(torch.randn(shape, generator=g).cuda())
        want = ops.maxpool_fwd(x, pd)
        got, sums = ops.maxpool_fwd_sums(x, pd)
        assert torch.equal(want.dense(), got.dense())
        y = got.dense().double()                                   # (n, d, h, w, c)
        ref = torch.stack([y.sum(dim=(1, 2, 3)), y.square().sum(dim=(1, 2, 3))], 1)
        assert torch.allclose(sums, ref, rtol=1e-6, atol=1e-6), (shape, (sums - ref).abs().max())
    x = ops.from_nchw(torch.randn((1, 8, 5, 6, 7), generator=g).cuda())
    got, sums = ops.maxpool_fwd_sums(x, 2)
    assert sums is None and torch.equal(got.dense(), ops.maxpool_fwd(x, 2).dense())


def test_graphed_train_step_matches_eager_steps(B):
    """GraphedTrainStep (forward + loss + backward + optimizer in one CUDA graph) against the same steps run eagerly
    from the same initial state: losses of 4 steps and the final parameters agree (the kernels and their order are
    the same; only atomics order differs), and new inputs are honoured.  SGD with momentum: linear in the gradient
    (Adam's first steps move every weight by lr * sign(g), which turns gradient noise around 0 into +-lr)."""
    def make():
        torch.manual_seed(71)
        net = B.UNet3D(1, 1, f_maps=[16, 32, 64], num_groups=4, layer_order="gcr").cuda().train()
        opt = torch.optim.SGD(net.parameters(), lr=0.05, momentum=0.9)
        crit = B.BCEDiceLoss(1.0, 1.0)
        return net, opt, (lambda x, y: crit(net(x), y))
    g = torch.Generator().manual_seed(72)
    xs = [torch.randn((2, 1, 32, 32, 32), generator=g).cuda() for _ in range(4)]
    ys = [(torch.rand((2, 1, 32, 32, 32), generator=g) > 0.5).float().cuda() for _ in range(4)]
    net_e, opt_e, fn_e = make()
    losses_e = []
    for x, y in zip(xs, ys):
        opt_e.zero_grad(set_to_none=True)
        l = fn_e(x, y)
        l.backward()
        opt_e.step()
        losses_e.append(l.item())
    net_g, opt_g, fn_g = make()
    state = {k: v.clone() for k, v in net_g.state_dict().items()}
    step = B.GraphedTrainStep(fn_g, opt_g, (xs[0], ys[0]))        # its warm-up steps move the parameters ...
    net_g.load_state_dict(state)                                  # ... so restore them and the optimizer state
    for st in opt_g.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()                                         # momentum buffers back to their initial state
    assert step.launches_per_step > 50
    losses_g = [step(x, y).item() for x, y in zip(xs, ys)]
    assert max(abs(a - b) for a, b in zip(losses_e, losses_g)) < 2e-3, (losses_e, losses_g)
    pe = torch.cat([p.detach().flatten() for p in net_e.parameters()])
    pg = torch.cat([p.detach().flatten() for p in net_g.parameters()])
    assert rel_l2(pg.cpu(), pe.cpu()) < 2e-3
    with pytest.raises(ValueError):
        step(xs[0][:1], ys[0][:1])


@pytest.mark.parametrize("shape,kd,ones", [((2, 1, 6, 9, 11), 3, True), ((1, 2, 5, 8, 8), 3, False),
                                            ((2, 3, 1, 12, 10), 1, False), ((1, 1, 1, 9, 16), 1, True),
                                            ((1, 4, 1, 8, 8), 1, False)])
def test_im2col_of_the_narrow_input_vs_torch(B, shape, kd, ones):
    """b200_im2col3: a[pos][tap * c + ci] = x[pos + tap][ci] (zero outside the volume), the optional synthesised
    ones column = 1 where the tap is inside, zero padding up to K' — bit-exact against a torch construction."""
    from mdeical_image_segmentation_b200 import ops
    n, c, d, h, w = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(shape, generator=g).bfloat16().float()
    xcl = ops.from_nchw(x.cuda(), pad=True)
    cin = c + (1 if ones else 0)
    taps = kd * 9
    kp = (taps * cin + 15) // 16 * 16
    a = ops.im2col(ops.CL(xcl.buf, c, 0, xcl.buf.shape[4] - c), kd, kp, ones=ones).dense().float().cpu()
    src = torch.cat([x, torch.ones((n, 1, d, h, w))], 1) if ones else x
    pd = kd // 2
    xp = F.pad(src, (1, 1, 1, 1, pd, pd))
    want = torch.zeros((n, d, h, w, kp))
    for t in range(taps):
        kw_, kh_, kd_ = t % 3, (t // 3) % 3, t // 9
        want[..., t * cin:(t + 1) * cin] = xp[:, :, kd_:kd_ + d, kh_:kh_ + h, kw_:kw_ + w].permute(0, 2, 3, 4, 1)
    assert torch.equal(a, want), (shape, (a - want).abs().max())
