"""ctypes binding of ``libb200unet.so`` (the C ABI declared in ``include/b200unet.h``).

The library is the product path: there is no CPU or PyTorch fallback.  ``lib()`` raises if the shared
object is missing and every wrapper raises ``RuntimeError(b200_strerror(rc))`` on a non-zero return.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libb200unet.so"
CSRC = _HERE / "csrc"

IMPL_AUTO, IMPL_TCGEN05, IMPL_DIRECT = 0, 1, 2


class ConvDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "n", "d", "h", "w", "c_in", "c_out", "kd", "kh", "kw",
        "in_ctot", "in_coff", "out_ctot", "out_coff", "relu", "impl")]


class ConvTDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "n", "h", "w", "c_in", "c_out", "in_ctot", "in_coff", "out_ctot", "out_coff", "impl")]


class DeconvDesc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "n", "d", "h", "w", "c_in", "c_out", "kd", "in_ctot", "in_coff", "impl")]


class RepackItem(C.Structure):
    _fields_ = [("w", C.c_void_p), ("wp_fwd", C.c_void_p), ("wp_bwd", C.c_void_p), ("c_out", C.c_int32),
                ("c_in", C.c_int32), ("c_in_pad", C.c_int32), ("ntaps", C.c_int32)]


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ctot", C.c_int32), ("coff", C.c_int32)]


_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_PC, _PT, _PD = C.POINTER(ConvDesc), C.POINTER(ConvTDesc), C.POINTER(DeconvDesc)

# name -> argtypes (restype is always int unless listed in _SPECIAL)
SIGNATURES = {
    "b200_conv_repack_weight": [_PC, _P, _P, _P, _P],
    "b200_conv_repack_weights": [_I, C.POINTER(RepackItem), _P],
    "b200_conv_fwd": [_PC, _P, _P, _P, _P, _P],
    "b200_conv_bwd_data": [_PC, _P, _P, _P, _P, _P],
    "b200_conv_bwd_weight": [_PC, _P, _P, _P, _P],
    "b200_conv_bwd_weight_bias": [_PC, _P, _P, _P, _P, _P],
    "b200_conv_fwd_stats": [_PC, _P, _P, _P, _P, _P, _P],
    "b200_conv_bwd_data_stats": [_PC, _P, _P, _P, _P, _I, _P, _I, _I, _P, _P],
    "b200_convt2x2_bwd_data_stats": [_PT, _P, _P, _P, _P, _I, _P, _P],
    "b200_gn_stats_finalize": [_I, _L, _I, _I, _F, _P, _P, _P],
    "b200_gn_bwd_fused": [_I, _L, _I, _I, View, _P, _P, View, View, _P, _P, _P, _P, _I, _P],
    "b200_gn_bwd_fused_up": [_I, _I, _I, _I, _I, _I, _I, _I, _I, _I, View, _P, _P, View, View, View, _P, _P, _P, _P,
                             _I, View, _P, _P],
    "b200_conv3d_c1_fwd": [_I, _I, _I, _I, _I, View, _P, _P, _I, View, _P, _P],
    "b200_gn_apply_up": [_I, _I, _I, _I, _I, _I, _I, _I, _I, _I, View, View, _P, _P, _P, _I, View, _P],
    "b200_bn_bwd_red_from_sums": [_I, _I, _P, _P, _P, _P, _P],
    "b200_convt2x2_repack_weight": [_PT, _P, _P, _P, _P],
    "b200_convt2x2_fwd": [_PT, _P, _P, _P, _P, _P],
    "b200_convt2x2_bwd_data": [_PT, _P, _P, _P, _P, _P],
    "b200_convt2x2_bwd_weight": [_PT, _P, _P, _P, _P],
    "b200_deconv3s2_repack_weight": [_PD, _P, _P, _P, _P],
    "b200_deconv_planes_resize_fwd": [_I, _I, _I, _I, _I, _I, _I, _I, _I, View, View, _P],
    "b200_deconv_planes_resize_bwd": [_I, _I, _I, _I, _I, _I, _I, _I, _I, View, View, _P],
    "b200_deconv3s2_fwd": [_PD, _P, _P, _P, _P],
    "b200_deconv3s2_bwd_data": [_PD, _P, _P, _P, _P],
    "b200_deconv3s2_bwd_weight": [_PD, _P, _P, _P, _P],
    "b200_maxpool2_fwd": [_I, _I, _I, _I, _I, _I, View, View, _P],
    "b200_maxpool2_fwd_sums": [_I, _I, _I, _I, _I, _I, View, View, _P, _P],
    "b200_maxpool2_bwd": [_I, _I, _I, _I, _I, _I, View, View, View, View, _I, _P],
    "b200_upsample_nearest_fwd": [_I, _I, _I, _I, _I, _I, _I, _I, View, View, _P],
    "b200_upsample_nearest_bwd": [_I, _I, _I, _I, _I, _I, _I, _I, View, View, _P],
    "b200_copy_channels": [_L, _I, View, View, _P],
    "b200_relu_bwd": [_L, _I, View, View, View, _P, _P],
    "b200_bias_grad": [_L, _I, View, _P, _P],
    "b200_maxpoolk_fwd": [_I, _I, _I, _I, _I, View, View, _P],
    "b200_maxpoolk_bwd": [_I, _I, _I, _I, _I, View, View, View, View, _P],
    "b200_bilinear2d_fwd": [_I, _I, _I, _I, _I, View, View, _P],
    "b200_bilinear2d_bwd": [_I, _I, _I, _I, _I, View, View, View, _P, _P],
    "b200_add": [_L, _I, View, View, _P],
    "b200_aug_flip_rot90": [_I, _I, _I, _I, _I, _I, _P, _P, _P],
    "b200_aug_affine_clip": [_I, _L, _P, _P, _P, _F, _F, _P, _P],
    "b200_aug_channel_stats": [_I, _L, _P, _P, _P],
    "b200_aug_gaussian_axis": [_I, _I, _I, _I, _F, _P, _P, _P],
    "b200_aug_map_coordinates": [_I, _I, _I, _I, _P, _P, _P, _P, _I, _P, _P],
    "b200_act_fwd": [_L, _I, _I, _F, View, View, View, _P],
    "b200_act_bwd": [_L, _I, _I, _F, View, View, View, _P],
    "b200_scale_mask": [_I, _L, _I, View, _P, _P, _F, View, _P],
    "b200_avgpool2_fwd": [_I, _I, _I, _I, _I, _I, View, View, _P],
    "b200_avgpool2_bwd": [_I, _I, _I, _I, _I, _I, View, View, View, View, _I, _P],
    "b200_add_relu": [_L, _I, View, View, View, _P],
    "b200_msssim_fwd": [_I, _I, _I, _P, _P, _P, _P, _P, _F, _P],
    "b200_msssim_bwd": [_I, _I, _I, _P, _P, _P, _P, _P, _F, _P],
    "b200_cgm_gate": [_I, _L, _I, View, _P, _P, _P, _F, _P, _P, _P],
    "b200_gate_sigmoid_fwd": [_I, _L, _P, _P, _P, _P],
    "b200_gate_sigmoid_bwd": [_I, _L, _P, _P, _P, _P, _P],
    "b200_patch_gather": [_I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "b200_patch_accumulate": [_I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _I, _P, _P, _P],
    "b200_prediction_finalize": [_I, _L, _P, _P, _P, _P],
    "b200_se_channel_sums": [_I, _L, _I, View, _P, _P],
    "b200_se_gate_fwd": [_I, _L, _I, View, _P, _P, View, _P],
    "b200_se_gate_bwd_reduce": [_I, _L, _I, View, _P, _P, View, _P, _P, _P],
    "b200_se_gate_bwd_apply": [_I, _L, _I, View, _P, _P, View, _P, View, View, _P],
    "b200_im2col3": [_I, _I, _I, _I, _I, _I, _I, View, _P, _P],
    "b200_nchw_f32_to_cl_bf16": [_I, _I, _L, _P, View, _P],
    "b200_cl_bf16_to_nchw_f32": [_I, _I, _L, View, _P, _P],
    "b200_gn_stats": [_I, _L, _I, _I, _F, View, _P, _P, _P],
    "b200_gn_apply": [_I, _L, _I, _I, View, _P, _P, _P, _I, View, _P],
    "b200_gn_bwd": [_I, _L, _I, _I, View, _P, _P, View, View, _P, _P, _P, _I, _P],
    "b200_bn_stats": [_L, _I, View, _P, _P],
    "b200_bn_apply": [_L, _I, View, _P, _P, _P, _P, _I, View, _P],
    "b200_bn_bwd_reduce": [_L, _I, View, View, View, _P, _P, _P, _P],
    "b200_bn_bwd_apply": [_L, _I, View, View, View, _P, _P, _P, _P, View, _I, _P],
    "b200_head_fwd": [_I, _L, _I, _I, View, _P, _P, _P, _P],
    "b200_head_bwd": [_I, _L, _I, _I, View, _P, _P, View, _P, _P, _I, _P],
    "b200_seg_sums": [_I, _I, _L, _P, _P, _I, _P, _P],
    "b200_seg_loss_bwd": [_I, _I, _L, _P, _P, _P, _I, _P, _P],
    "b200_ce_fwd": [_I, _I, _L, _P, _P, _P, _P],
    "b200_ce_bwd": [_I, _I, _L, _P, _P, _P, _P, _P],
    "b200_metric_prob_sum": [_L, _P, _P, _P],
    "b200_metric_counts": [_I, _L, _P, _P, _P, _I, _P, _P],
    "b200_metric_argmax_counts": [_I, _I, _L, _P, _P, _P, _P],
}
_SPECIAL = {
    "b200_version": ([], C.c_int),
    "b200_strerror": ([C.c_int], C.c_char_p),
    "b200_last_impl": ([], C.c_int),
    "b200_set_tuning": ([C.c_int, C.c_int], C.c_int),
    "b200_msssim_pyramid_floats": ([_I, _I, _I], C.c_int64),
}
ALL_SYMBOLS = sorted(list(SIGNATURES) + list(_SPECIAL))

_lib = None


def build(force: bool = False) -> Path:
    """Compile the CUDA sources in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if force:
        subprocess.run(["make", "-C", str(CSRC), "clean"], check=True, capture_output=True)
    r = subprocess.run(["make", "-C", str(CSRC), "-j", str(os.cpu_count() or 4)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libb200unet.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the U-Net hot path)")
        l = C.CDLL(str(LIB_PATH))
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        for name, (argtypes, restype) in _SPECIAL.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = l
    return _lib


def strerror(rc: int) -> str:
    return lib().b200_strerror(rc).decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"{what or 'b200 call'} failed ({rc}): {strerror(rc)}")


# kernels launched per successful C-ABI call (for bench.py's `gpu_launches` accounting)
LAUNCHES_PER_CALL = {"b200_gn_stats": 2, "b200_gn_bwd": 3, "b200_gn_bwd_fused": 2, "b200_deconv3s2_fwd": 8,
                     "b200_gn_bwd_fused_up": 3, "b200_gn_apply_up": 2}
launch_count = 0
# optional profiler hook: fn(name, args) -> context manager, installed by bench.py for per-kernel timing
profile_hook = None


def call(name: str, *args) -> None:
    global launch_count
    if profile_hook is not None:
        with profile_hook(name, args):
            check(getattr(lib(), name)(*args), name)
    else:
        check(getattr(lib(), name)(*args), name)
    launch_count += LAUNCHES_PER_CALL.get(name, 1)
