"""CPU: the oracle restatement replayed against the round-2 golden fixtures (tests/golden/blocks_r2.npz,
resunet_r2.npz — outputs of the UNMODIFIED reference, made by oracle/make_golden_r2.py): remaining layer_order
letters, ResNetBlock orders, ResNetBlockSE modes, the SE layers, Decoder / Encoder stand-alone forwards,
ResidualUNet2D, ELU / LeakyReLU residual networks, get_model; plus the LazyPredictor restatement."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import predictor as OP
from oracle import unet as O


def cases(fix):
    return sorted({k.split("/")[0] for k in fix if "/" in k})


def case_of(fix, case):
    sd = {k[len(case) + 4:]: torch.from_numpy(v.copy()) for k, v in fix.items() if k.startswith(case + "/sd/")}
    ins = {k.split("/")[2]: torch.from_numpy(v.copy()) for k, v in fix.items() if k.startswith(case + "/in/")}
    dins = {k.split("/")[2]: torch.from_numpy(v.copy()) for k, v in fix.items() if k.startswith(case + "/din/")}
    grads = {k[len(case) + 6:]: torch.from_numpy(v.copy()) for k, v in fix.items() if k.startswith(case + "/grad/")}
    return sd, ins, dins, grads


def oracle_block(case, sd, ins, training=True):
    """The oracle call that restates the reference module of a blocks_r2 case."""
    kind, _, tail = case.partition("_")
    is3d = "2d" not in kind
    if kind.startswith("sc"):
        order = tail.replace("eval_", "")
        return O.single_conv(sd, "", ins["x"], order, 4, is3d, training and "eval" not in tail, dropout_prob=0.3)
    if kind.startswith("rbse"):
        return O.resnet_block(sd, "", ins["x"], "gcr", 4, True, training)
    if kind.startswith("rb"):
        order = tail.split("_")[0]
        return O.resnet_block(sd, "", ins["x"], order, 4, is3d, training)
    if kind == "se":
        return O.se_layer(sd, "", ins["x"])
    if kind.startswith("dec"):
        order = {"deconv": "cbr", "nearest": "cr" if not is3d else "gcr", "res": "gcr"}[tail.split("_")[0]]
        return O.decoder_forward(sd, "", ins["skip"], ins["x"], order=order, num_groups=4, is3d=is3d,
                                 training=training, residual=tail == "res")
    if kind.startswith("enc"):
        order = "cr" if not is3d else "gcr"
        return O.encoder_forward(sd, "", ins["x"], order=order, num_groups=4, is3d=is3d, training=training,
                                 pool="avg" if "avg" in tail else "max", residual="res" in tail)
    raise KeyError(case)


BLOCKS = load_golden("blocks_r2")
NETS = load_golden("resunet_r2")


@pytest.mark.parametrize("case", cases(BLOCKS))
def test_oracle_blocks_vs_reference(case):
    sd, ins, dins, grads = case_of(BLOCKS, case)
    for k, v in sd.items():
        if v.is_floating_point() and "running_" not in k:
            v.requires_grad_(True)
    for v in ins.values():
        v.requires_grad_(True)
    y = oracle_block(case, sd, ins)
    assert rel_l2(y.detach(), BLOCKS[case + "/y"]) < 2e-5, case
    (y * torch.from_numpy(BLOCKS[case + "/r"])).sum().backward()
    for k, v in ins.items():
        assert rel_l2(v.grad, dins[k]) < 1e-4, (case, k)
    for k, g in grads.items():
        got = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        assert (got - g).norm() <= 1e-4 * g.norm() + 1e-5, (case, k)


@pytest.mark.parametrize("case,kw", [("resunet2d_gcr", dict(num_levels=3, layer_order="gcr", num_groups=4, is3d=False)),
                                     ("resunet3d_cge", dict(num_levels=3, layer_order="cge", num_groups=4)),
                                     ("resunet3d_gcl", dict(num_levels=2, layer_order="gcl", num_groups=4))])
def test_oracle_residual_networks_vs_reference(case, kw):
    sd = {k[len(case) + 4:]: torch.from_numpy(v.copy()).requires_grad_("running_" not in k and v.dtype.kind == "f")
          for k, v in NETS.items() if k.startswith(case + "/sd/")}
    x, t = torch.from_numpy(NETS[case + "/x"]), torch.from_numpy(NETS[case + "/t"])
    logits = O.residual_unet3d_forward(sd, x, training=True, **kw)
    assert rel_l2(logits.detach(), NETS[case + "/logits"]) < 2e-5
    loss = O.bce_dice_loss(logits, t)
    assert abs(loss.item() - float(NETS[case + "/loss"])) < 1e-5
    loss.backward()
    for k, v in sd.items():
        g = torch.from_numpy(NETS[f"{case}/grad/{k}"])
        assert (v.grad - g).norm() <= 2e-4 * g.norm() + 1e-6, k
    with torch.no_grad():
        ev = O.residual_unet3d_forward(sd, x, training=False, **kw)
    assert rel_l2(ev, NETS[case + "/logits_eval"]) < 2e-5


def test_get_model_factory_matches_reference():
    """model/unet3d/model.py:362-364: the class is picked by config['name'] and receives the whole dict."""
    import mdeical_image_segmentation_b200 as B
    for entry in NETS["get_model"]:
        name, want = str(entry).split("=")
        cls_name, n_keys = want.split(":")
        m = B.get_model(dict(name=name, in_channels=1, out_channels=1, f_maps=[8, 16], num_groups=4))
        assert type(m).__name__ == cls_name and len(m.state_dict()) == int(n_keys), entry
    with pytest.raises(Exception):
        B.get_model(dict(name="NoSuchNet", in_channels=1, out_channels=1))


def test_state_dict_keys_of_new_blocks_are_the_reference_ones():
    import mdeical_image_segmentation_b200 as B
    makers = {
        "sc3d_cge": lambda: B.SingleConv(4, 8, order="cge", num_groups=4),
        "sc3d_eval_cgrD": lambda: B.SingleConv(4, 8, order="cgrD", num_groups=4, dropout_prob=0.3),
        "rb3d_cge": lambda: B.ResNetBlock(4, 8, order="cge", num_groups=4),
        "rb2d_gcr": lambda: B.ResNetBlock(4, 8, order="gcr", num_groups=4, is3d=False),
        "rbse_cse": lambda: B.ResNetBlockSE(4, 8, order="gcr", num_groups=4, se_module="cse"),
        "rbse_sse": lambda: B.ResNetBlockSE(4, 8, order="gcr", num_groups=4, se_module="sse"),
        "rbse_scse": lambda: B.ResNetBlockSE(4, 8, order="gcr", num_groups=4, se_module="scse"),
        "se_cse": lambda: B.ChannelSELayer3D(8, 2),
        "se_sse": lambda: B.SpatialSELayer3D(8),
        "dec3d_res": lambda: B.Decoder(16, 8, basic_module=B.ResNetBlock, conv_layer_order="gcr", num_groups=4),
        "dec3d_deconv": lambda: B.Decoder(16, 8, conv_layer_order="cbr", upsample="deconv"),
        "enc3d_avg": lambda: B.Encoder(4, 8, pool_type="avg", conv_layer_order="gcr", num_groups=4),
    }
    for case, make in makers.items():
        sd, _, _, _ = case_of(BLOCKS, case)
        mine = make().state_dict()
        assert set(mine) == set(sd), (case, set(mine) ^ set(sd))
        for k in sd:
            assert tuple(mine[k].shape) == tuple(sd[k].shape), (case, k)
    net = B.ResidualUNet2D(2, 2, f_maps=[8, 16, 32], num_groups=4)
    want = {k[len("resunet2d_gcr/sd/"):] for k in NETS if k.startswith("resunet2d_gcr/sd/")}
    assert set(net.state_dict()) == want


def test_lazy_predictor_restatement_equals_standard():
    """predictor.py:171-223: block-wise normalisation with the mask reset gives exactly the in-memory result,
    also when 1/27-volume blocks overlap (sizes not divisible by 3) and with a segmentation output."""
    rng = np.random.default_rng(0)
    raw = rng.standard_normal((2, 11, 13, 10)).astype(np.float32)

    def model_fn(batch):
        b = np.asarray(batch, dtype=np.float32)
        pos = np.arange(b.shape[-1], dtype=np.float32)
        return np.stack([np.tanh(b.sum(1)) + 0.01 * pos, np.cos(b[:, 0]) - 0.02 * pos,
                         0.5 * b[:, -1]], axis=1).astype(np.float32)

    for seg in (False, True):
        a = OP.standard_predict(model_fn, raw, (6, 8, 6), (3, 5, 4), (1, 2, 1), out_channels=3, save_segmentation=seg,
                                batch_size=2)
        b = OP.lazy_predict(model_fn, raw, (6, 8, 6), (3, 5, 4), (1, 2, 1), out_channels=3, save_segmentation=seg,
                            batch_size=2)
        assert a.dtype == b.dtype and np.array_equal(a, b)


# ----------------------------------------------------------------------------------- augmentation
AUG = load_golden("augment_r2")


def test_oracle_augmentations_vs_reference():
    """oracle/augment.py against the outputs of the reference's own transforms (augment/unet3d_augment/transforms.py)
    for the same np.random.RandomState seeds; its gaussian_filter / map_coordinates restatements against scipy."""
    from scipy import ndimage
    from oracle import augment as OA
    m4, m3, lab = AUG["m4"], AUG["m3"], AUG["lab"]
    for seed in range(6):
        assert np.array_equal(OA.random_flip(m4, np.random.RandomState(seed)), AUG[f"flip/{seed}"])
        assert np.array_equal(OA.random_rotate90(m4, np.random.RandomState(seed)), AUG[f"rot/{seed}"])
        assert np.array_equal(OA.random_rotate90(m3, np.random.RandomState(seed)), AUG[f"rot3/{seed}"])
        assert np.allclose(OA.random_contrast(m4, np.random.RandomState(seed), execution_probability=0.7),
                           AUG[f"contrast/{seed}"], atol=1e-7)
    src = {"m4": m4, "m3": m3, "lab": lab}
    for key in [k for k in AUG if k.startswith("elastic/")]:
        seed, order, name = key.split("/")[1].split("_")
        got = OA.elastic_deformation(src[name], np.random.RandomState(int(seed)), int(order), alpha=15, sigma=2,
                                     execution_probability=1.0)
        assert np.allclose(got, AUG[key], atol=2e-6), key
    got = OA.elastic_deformation(m3, np.random.RandomState(4), 1, alpha=15, sigma=2, execution_probability=1.0,
                                 apply_3d=False)
    assert np.allclose(got, AUG["elastic2d/4_1_m3"], atol=2e-6)
    assert np.allclose(OA.standardize(m4), AUG["standardize/global"], atol=1e-6)
    assert np.allclose(OA.standardize(m4, channelwise=True), AUG["standardize/channelwise"], atol=1e-6)
    assert np.allclose(OA.normalize(m4 * 3), AUG["normalize/default"], atol=1e-6)
    assert np.allclose(OA.normalize(m4 * 3, norm01=True), AUG["normalize/norm01"], atol=1e-6)
    rs = np.random.RandomState(1)
    field = rs.randn(6, 9, 10)
    assert np.abs(OA.gaussian_filter(field, 2.0) - ndimage.gaussian_filter(field, 2.0, mode="reflect")).max() < 1e-12
    coords = [rs.uniform(-12, 20, size=(6, 9, 10)) for _ in range(3)]
    for order in (0, 1):
        ref = ndimage.map_coordinates(m3[:6, :8, :10].copy(), coords, order=order, mode="reflect")
        assert np.abs(OA.map_coordinates(m3[:6, :8, :10].copy(), coords, order) - ref).max() < 1e-6
