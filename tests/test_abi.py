"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU or libcuda, exports
every symbol include/b200unet.h declares, and the ctypes binding covers exactly that set."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from mdeical_image_segmentation_b200 import _lib

HEADER = os.path.join(ROOT, "include", "b200unet.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_loads():
    _lib.build()
    lib = _lib.lib()
    assert lib.b200_version() >= 100
    assert b"unsupported" in lib.b200_strerror(-2)


def test_every_declared_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 40
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in b200unet.h but not exported"
    assert sorted(_lib.ALL_SYMBOLS) == syms, set(_lib.ALL_SYMBOLS) ^ set(syms)


def test_bad_arguments_return_error_codes_not_exceptions():
    lib = _lib.lib()
    d = _lib.ConvDesc(1, 1, 4, 4, 8, 8, 5, 5, 5, 8, 0, 8, 0, 0, 0)   # 5x5x5 kernel: unsupported
    assert lib.b200_conv_fwd(ctypes.byref(d), None, None, None, None, None) == -2
    d = _lib.ConvDesc(0, 1, 4, 4, 8, 8, 3, 3, 3, 8, 0, 8, 0, 0, 0)   # empty batch: bad arg
    assert lib.b200_conv_fwd(ctypes.byref(d), None, None, None, None, None) == -1
    d = _lib.ConvDesc(1, 1, 4, 4, 8, 8, 1, 3, 3, 8, 4, 8, 0, 0, 0)   # slice overruns its buffer
    assert lib.b200_conv_fwd(ctypes.byref(d), None, None, None, None, None) == -1


def test_no_cpu_fallback():
    import mdeical_image_segmentation_b200 as B
    net = B.UNet(1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 1, 16, 16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.DiceLoss()(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 1, 4, 4, 4))


def test_state_dict_layouts_are_the_reference_ones():
    import mdeical_image_segmentation_b200 as B
    from conftest import load_golden
    fix = load_golden("unet3d_cbr_deconv")
    want = sorted(k[4:] for k in fix if k.startswith("sd__"))
    net = B.UNet3D(2, 2, f_maps=[8, 16, 32], num_groups=4, layer_order="cbr", upsample="deconv")
    assert sorted(net.state_dict()) == want
    fix = load_golden("unet3d_gcr_nearest")
    want = sorted(k[4:] for k in fix if k.startswith("sd__"))
    net = B.UNet3D(2, 2, f_maps=[8, 16, 32], num_groups=4)
    assert sorted(net.state_dict()) == want
    sd2 = B.UNetModel(B.UNetConfig(3, 4)).state_dict()
    assert len(sd2) == 46 and "unet.up_sample.3.up.weight" in sd2 and "unet.final_conv.bias" in sd2
    assert tuple(sd2["unet.up_sample.0.up.weight"].shape) == (1024, 512, 2, 2)


def test_msssim_pyramid_layout_host_side():
    """b200_msssim_pyramid_floats is pure host arithmetic: 5 scales, ceil-halving, 4-float alignment, and the
    package's size assertion (smaller side > 160) as an error code — must agree with ops.msssim_valid_counts."""
    from mdeical_image_segmentation_b200 import _lib, ops
    lib = _lib.lib()
    for n, h, w in ((2, 176, 192), (1, 161, 173), (3, 512, 512)):
        total, hh, ww, counts = 0, h, w, []
        for _ in range(5):
            total += (n * hh * ww + 3) // 4 * 4
            counts.append((hh - 10) * (ww - 10))
            hh, ww = (hh + 1) // 2, (ww + 1) // 2
        assert lib.b200_msssim_pyramid_floats(n, h, w) == total
        assert ops.msssim_valid_counts(h, w) == counts and min(counts) >= 1
    assert lib.b200_msssim_pyramid_floats(1, 160, 400) < 0 and lib.b200_msssim_pyramid_floats(1, 400, 100) < 0


def test_new_entry_points_validate_before_touching_the_device():
    """The UNet 3+ / MS-SSIM / SE / predictor entry points reject unsupported arguments with error codes before any
    CUDA call (so this runs without a GPU), and the widened networks keep the no-CPU-fallback rule."""
    import mdeical_image_segmentation_b200 as B
    lib = _lib.lib()
    V = _lib.View
    ok16 = V(ctypes.c_void_p(4096), 16, 0)            # never dereferenced: validation fails first
    odd = V(ctypes.c_void_p(4096), 12, 0)             # channel total not a multiple of 8
    assert lib.b200_bilinear2d_fwd(1, 4, 4, 16, 2, odd, ok16, None) == -2
    assert lib.b200_bilinear2d_fwd(1, 4, 4, 16, 64, ok16, ok16, None) == -2          # scale > 32
    assert lib.b200_bilinear2d_bwd(1, 4, 4, 16, 2, ok16, V(None, 0, 0), ok16, None, None) == -2   # no scratch
    assert lib.b200_maxpoolk_fwd(1, 6, 8, 16, 4, ok16, ok16, None) == -2              # 6 % 4 != 0
    assert lib.b200_add_relu(4, 12, ok16, ok16, ok16, None) == -2
    assert lib.b200_msssim_fwd(1, 160, 400, None, None, None, None, None, ctypes.c_float(1.0), None) == -2
    assert lib.b200_msssim_bwd(1, 100, 100, None, None, None, None, None, ctypes.c_float(1.0), None) == -2
    big = V(ctypes.c_void_p(4096), 2048, 0)
    assert lib.b200_se_channel_sums(1, 8, 2048, big, ctypes.c_void_p(4096), None) == -2   # c > 1024
    assert lib.b200_se_gate_fwd(1, 8, 16, ok16, None, None, ok16, None) == -2             # no gates
    assert lib.b200_cgm_gate(1, 4, 12, odd, None, None, None, ctypes.c_float(1.0), None, None, None) == -2
    assert lib.b200_patch_gather(1, 8, 8, 8, None, 0, 0, 0, 0, 0, 0, 4, 4, 4, None, None) == -1   # null buffers
    assert lib.b200_patch_gather(1, 8, 8, 8, ctypes.c_void_p(4096), 0, 0, 0, 8, 0, 0, 4, 4, 4,
                                 ctypes.c_void_p(4096), None) == -1                               # halo >= size
    assert lib.b200_patch_accumulate(1, 8, 8, 8, 4, 4, 4, 0, 0, 0, ctypes.c_void_p(4096), 1, -1, 6, 0, 0,
                                     ctypes.c_void_p(4096), ctypes.c_void_p(4096), None) == -1    # patch leaves the map
    assert lib.b200_prediction_finalize(0, 8, None, None, None, None) == -1
    for make, shape in ((lambda: B.UNet_3Plus(3, 1), (1, 3, 32, 32)),
                        (lambda: B.ResidualUNet3D(1, 1, f_maps=[8, 16]), (1, 1, 8, 8, 8))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            make()(torch.zeros(shape))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.MSSSIMLoss()(torch.zeros(1, 1, 176, 176), torch.zeros(1, 1, 176, 176))
    from mdeical_image_segmentation_b200.predictor import StandardPredictor
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        StandardPredictor(torch.nn.Conv3d(1, 1, 1), None, 1).predict_volume(torch.zeros(8, 64, 64), (8, 64, 64), (8, 64, 64))
