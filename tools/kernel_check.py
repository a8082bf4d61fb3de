"""On-GPU bring-up check of the raw kernels against stock PyTorch (fp32 math on bf16-rounded data).

Run as `python tools/kernel_check.py --group NAME` (one process per group so that a faulting kernel
cannot poison the others); `tools/run_kernel_checks.sh` drives all groups under `timeout`.
Not part of the product path.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdeical_image_segmentation_b200 import ops  # noqa: E402
from mdeical_image_segmentation_b200.ops import CL  # noqa: E402

dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rnd_cl(n, d, h, w, c, ctot=None, coff=0, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    ctot = ctot or c
    buf = (torch.randn((n, d, h, w, ctot), generator=g) * scale).to(torch.bfloat16).to(dev)
    return CL(buf, c, coff)


def relerr(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()


def report(name, a, b, tol=1e-2):
    r, m = relerr(a, b)
    ok = r < tol and r == r
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel_l2={r:.3e} max_abs={m:.3e}", flush=True)
    return ok


def ref_conv(x: CL, w, bias, relu):
    xin = x.to_nchw()
    wq = w.to(torch.bfloat16).float()
    if w.dim() == 4:
        y = F.conv2d(xin[:, :, 0], wq, bias, padding=w.shape[-1] // 2).unsqueeze(2)
    else:
        y = F.conv3d(xin, wq, bias, padding=tuple(k // 2 for k in w.shape[2:]))
    return F.relu(y) if relu else y


CONV_CASES = [
    # name, n,d,h,w, cin,cout, 3d?, in(ctot,coff), out(ctot,coff), bias, relu
    ("2d_64_64", 2, 1, 32, 32, 64, 64, False, None, None, True, True),
    ("2d_partial_128_256", 1, 1, 24, 40, 128, 256, False, None, None, True, False),
    ("3d_64_128", 1, 16, 16, 16, 64, 128, True, None, None, False, True),
    ("3d_k32_32_64", 2, 8, 8, 8, 32, 64, True, None, None, False, False),
    ("2d_slices", 1, 1, 16, 16, 64, 64, False, (192, 64), (128, 64), True, True),
    ("2d_16_16", 1, 1, 16, 16, 16, 16, False, None, None, True, True),
    ("3d_odd_48_32", 1, 5, 6, 7, 48, 32, True, None, None, False, True),
    ("2d_512_wide", 1, 1, 8, 8, 512, 512, False, None, None, True, True),
]


def group_conv_fwd(impls):
    ok = True
    for (name, n, d, h, w, cin, cout, is3d, ins, outs, has_bias, relu) in CONV_CASES:
        ictot, icoff = ins or (cin, 0)
        octot, ocoff = outs or (cout, 0)
        x = rnd_cl(n, d, h, w, cin, ictot, icoff, seed=1)
        g = torch.Generator().manual_seed(2)
        wshape = (cout, cin, 3, 3, 3) if is3d else (cout, cin, 3, 3)
        wt = (torch.randn(wshape, generator=g) / (cin * 9) ** 0.5).to(dev)
        bias = torch.randn(cout, generator=g).to(dev) if has_bias else None
        fwd, bwd, k = ops.conv_repack_weight(wt)
        ref = ref_conv(x, wt, bias, relu)
        for impl in impls:
            outbuf = torch.full((n, d, h, w, octot), 7.0, dtype=torch.bfloat16, device=dev)
            out = CL(outbuf, cout, ocoff)
            try:
                ops.conv_fwd(x, fwd, k, bias, relu, out=out, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"conv_fwd[{name}] impl={impl}", out.to_nchw(), ref)
                if octot != cout:  # untouched channels stay untouched
                    other = torch.cat([outbuf[..., :ocoff], outbuf[..., ocoff + cout:]], -1)
                    assert (other == 7.0).all(), "wrote outside the channel slice"
            except Exception as e:  # noqa
                ok = False
                print(f"[BAD] conv_fwd[{name}] impl={impl}: {e}", flush=True)
    return ok


def group_conv_bwd(impls):
    ok = True
    for (name, n, d, h, w, cin, cout, is3d, ins, outs, has_bias, relu) in CONV_CASES:
        x = rnd_cl(n, d, h, w, cin, seed=3)
        dy = rnd_cl(n, d, h, w, cout, seed=4)
        g = torch.Generator().manual_seed(5)
        wshape = (cout, cin, 3, 3, 3) if is3d else (cout, cin, 3, 3)
        wt = (torch.randn(wshape, generator=g) / (cin * 9) ** 0.5).to(dev)
        fwd, bwd, k = ops.conv_repack_weight(wt)
        xin = x.to_nchw().requires_grad_(True)
        wq = wt.to(torch.bfloat16).float().requires_grad_(True)
        if is3d:
            y = F.conv3d(xin, wq, None, padding=1)
            y.backward(dy.to_nchw())
        else:
            y = F.conv2d(xin[:, :, 0], wq, None, padding=1)
            y.backward(dy.to_nchw()[:, :, 0])
        dx_ref = xin.grad
        dw_ref = wq.grad
        mask = rnd_cl(n, d, h, w, cin, seed=6)
        for impl in impls:
            try:
                dx = ops.conv_bwd_data(dy, bwd, k, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"conv_dgrad[{name}] impl={impl}", dx.to_nchw(), dx_ref)
                dxm = ops.conv_bwd_data(dy, bwd, k, mask=mask, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"conv_dgrad_mask[{name}] impl={impl}", dxm.to_nchw(),
                             dx_ref * (mask.to_nchw() > 0))
                dw = ops.conv_bwd_weight(x, dy, k, wt.shape, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"conv_wgrad[{name}] impl={impl}", dw, dw_ref)
            except Exception as e:  # noqa
                ok = False
                print(f"[BAD] conv_bwd[{name}] impl={impl}: {e}", flush=True)
    return ok


def group_convt(impls):
    ok = True
    for (n, h, w, cin, cout) in [(2, 16, 16, 128, 64), (1, 12, 20, 256, 128), (1, 8, 8, 64, 32),
                                 (1, 4, 4, 1024, 512)]:
        x = rnd_cl(n, 1, h, w, cin, seed=7)
        dy = rnd_cl(n, 1, 2 * h, 2 * w, cout, seed=8)
        g = torch.Generator().manual_seed(9)
        wt = (torch.randn((cin, cout, 2, 2), generator=g) / cin ** 0.5).to(dev)
        bias = torch.randn(cout, generator=g).to(dev)
        fwd, bwd = ops.convt_repack_weight(wt)
        xin = x.to_nchw()[:, :, 0].requires_grad_(True)
        wq = wt.to(torch.bfloat16).float().requires_grad_(True)
        y = F.conv_transpose2d(xin, wq, bias, stride=2)
        y.backward(dy.to_nchw()[:, :, 0])
        for impl in impls:
            try:
                out = ops.convt_fwd(x, fwd, bias, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"convt_fwd[{n},{h},{w},{cin}->{cout}] impl={impl}", out.to_nchw()[:, :, 0], y)
                dx = ops.convt_bwd_data(dy, bwd, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"convt_dgrad[{cin}->{cout}] impl={impl}", dx.to_nchw()[:, :, 0], xin.grad)
                dw = ops.convt_bwd_weight(x, dy, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"convt_wgrad[{cin}->{cout}] impl={impl}", dw, wq.grad)
            except Exception as e:  # noqa
                ok = False
                print(f"[BAD] convt[{cin}->{cout}] impl={impl}: {e}", flush=True)
    return ok


def group_deconv(impls):
    ok = True
    for (n, d, h, w, cin, cout, is3d) in [(1, 4, 6, 8, 32, 16, True), (2, 8, 8, 8, 128, 64, True),
                                          (1, 1, 12, 20, 64, 32, False), (1, 16, 16, 16, 64, 64, True)]:
        x = rnd_cl(n, d, h, w, cin, seed=11)
        g = torch.Generator().manual_seed(12)
        ks = (3, 3, 3) if is3d else (3, 3)
        wt = (torch.randn((cin, cout) + ks, generator=g) / (cin * 4) ** 0.5).to(dev)
        size = (2 * d, 2 * h, 2 * w) if is3d else (1, 2 * h, 2 * w)
        dy = rnd_cl(n, size[0], size[1], size[2], cout, seed=13)
        xin = x.to_nchw().requires_grad_(True)
        wq = wt.to(torch.bfloat16).float().requires_grad_(True)
        if is3d:
            y = F.interpolate(F.conv_transpose3d(xin, wq, None, stride=2, padding=1), size=size)
            y.backward(dy.to_nchw())
            yref = y
        else:
            y = F.interpolate(F.conv_transpose2d(xin[:, :, 0], wq, None, stride=2, padding=1), size=size[1:])
            y.backward(dy.to_nchw()[:, :, 0])
            yref = y.unsqueeze(2)
        fwd, bwd, kd = ops.deconv_repack_weight(wt)
        for impl in impls:
            try:
                planes = ops.deconv_fwd(x, fwd, kd, impl=impl)
                out = ops.deconv_resize_fwd(planes, n, kd, size)
                torch.cuda.synchronize()
                ok &= report(f"deconv_fwd[{n},{d},{h},{w},{cin}->{cout}] impl={impl}", out.to_nchw(), yref)
                dpl = ops.deconv_resize_bwd(dy, x.dims, kd)
                dx = ops.deconv_bwd_data(dpl, bwd, x.dims, kd, impl=impl)
                dw = ops.deconv_bwd_weight(x, dpl, wt.shape, kd, impl=impl)
                torch.cuda.synchronize()
                ok &= report(f"deconv_dgrad[{cin}->{cout}] impl={impl}", dx.to_nchw(), xin.grad)
                ok &= report(f"deconv_wgrad[{cin}->{cout}] impl={impl}", dw, wq.grad, tol=2e-2)
            except Exception as e:  # noqa
                ok = False
                print(f"[BAD] deconv[{cin}->{cout}] impl={impl}: {e}", flush=True)
    return ok


def group_perf():
    """First timing of the tcgen05 kernels on the headline layer shapes."""
    shapes = [
        ("cfg3 dec2.conv1 192->64 128^3 b2", 2, 128, 128, 128, 192, 64, True),
        ("cfg3 dec2.conv2 64->64 128^3 b2", 2, 128, 128, 128, 64, 64, True),
        ("cfg3 dec1.conv1 384->128 64^3 b2", 2, 64, 64, 64, 384, 128, True),
        ("cfg3 dec0.conv1 768->256 32^3 b2", 2, 32, 32, 32, 768, 256, True),
        ("cfg2 up3.first 128->64 512^2 b16", 16, 1, 512, 512, 128, 64, False),
        ("cfg2 down1.second 128->128 256^2 b16", 16, 1, 256, 256, 128, 128, False),
        ("cfg2 mid.second 1024->1024 32^2 b64", 64, 1, 32, 32, 1024, 1024, False),
        ("cfg2 up0.first 1024->512 64^2 b16", 16, 1, 64, 64, 1024, 512, False),
    ]
    for (name, n, d, h, w, cin, cout, is3d) in shapes:
        x = CL(torch.randn((n, d, h, w, cin), device=dev).to(torch.bfloat16))
        dy = CL(torch.randn((n, d, h, w, cout), device=dev).to(torch.bfloat16))
        wshape = (cout, cin, 3, 3, 3) if is3d else (cout, cin, 3, 3)
        wt = torch.randn(wshape, device=dev) * 0.02
        fwd, bwd, k = ops.conv_repack_weight(wt)
        flops = 2.0 * n * d * h * w * cin * cout * (27 if is3d else 9)
        out = CL.empty(n, d, h, w, cout, dev)
        dx = CL.empty(n, d, h, w, cin, dev)
        from mdeical_image_segmentation_b200 import _lib
        def _nohalo(f):
            def g():
                _lib.lib().b200_set_tuning(0, 0)
                try:
                    return f()
                finally:
                    _lib.lib().b200_set_tuning(0, 1)
            return g
        def _nohalo_w(f):
            def g():
                _lib.lib().b200_set_tuning(1, 0)
                try:
                    return f()
                finally:
                    _lib.lib().b200_set_tuning(1, 1)
            return g
        for label, fn in [
            ("fwd", lambda: ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1)),
            ("fwd-nohalo", _nohalo(lambda: ops.conv_fwd(x, fwd, k, None, True, out=out, impl=1))),
            ("dgrad", lambda: ops.conv_bwd_data(dy, bwd, k, out=dx, impl=1)),
            ("dgrad-nohalo", _nohalo(lambda: ops.conv_bwd_data(dy, bwd, k, out=dx, impl=1))),
            ("wgrad", lambda: ops.conv_bwd_weight(x, dy, k, wt.shape, impl=1)),
            ("wgrad-nohalo", _nohalo_w(lambda: ops.conv_bwd_weight(x, dy, k, wt.shape, impl=1))),
        ]:
            try:
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                print(f"[PERF] {name} {label}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
            except Exception as e:  # noqa
                print(f"[BAD] perf {name} {label}: {e}", flush=True)
        # cuDNN incumbent (bf16 channels_last)
        try:
            xt = x.dense().permute(0, 4, 1, 2, 3)
            wb = wt.to(torch.bfloat16)
            if is3d:
                wb = wb.to(memory_format=torch.channels_last_3d)
                f = lambda: F.conv3d(xt, wb, padding=1)  # noqa
            else:
                xt2 = xt[:, :, 0]
                wb = wb.to(memory_format=torch.channels_last)
                f = lambda: F.conv2d(xt2, wb, padding=1)  # noqa
            torch.backends.cudnn.benchmark = True
            for _ in range(3):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                f()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"[PERF] {name} cudnn-fwd: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
        except Exception as e:  # noqa
            print(f"[BAD] cudnn {name}: {e}", flush=True)


def group_wgrad_perf():
    """Weight-gradient kernels on the headline layer shapes: narrow (N = 64, five tap pairs) vs wide
    (N = 128, four tap pairs + centre-tap GEMM) variants."""
    from mdeical_image_segmentation_b200 import _lib
    shapes = [
        ("3D 384->128 64^3 b2", 2, 64, 64, 64, 384, 128, True),
        ("3D 128->128 64^3 b2", 2, 64, 64, 64, 128, 128, True),
        ("3D 768->256 32^3 b2", 2, 32, 32, 32, 768, 256, True),
        ("3D 256->256 32^3 b2", 2, 32, 32, 32, 256, 256, True),
        ("3D 64->128 64^3 b2", 2, 64, 64, 64, 64, 128, True),
        ("3D 256->512 16^3 b2", 2, 16, 16, 16, 256, 512, True),
        ("2D 64->128 256^2 b64", 64, 1, 256, 256, 64, 128, False),
        ("2D 128->128 256^2 b64", 64, 1, 256, 256, 128, 128, False),
        ("2D 256->128 256^2 b64", 64, 1, 256, 256, 256, 128, False),
        ("2D 256->256 128^2 b64", 64, 1, 128, 128, 256, 256, False),
        ("2D 512->256 128^2 b64", 64, 1, 128, 128, 512, 256, False),
        ("2D 512->512 64^2 b64", 64, 1, 64, 64, 512, 512, False),
        ("2D 1024->512 64^2 b64", 64, 1, 64, 64, 1024, 512, False),
        ("2D 1024->1024 32^2 b64", 64, 1, 32, 32, 1024, 1024, False),
    ]
    for (name, n, d, h, w, cin, cout, is3d) in shapes:
        x = CL(torch.randn((n, d, h, w, cin), device=dev).to(torch.bfloat16))
        dy = CL(torch.randn((n, d, h, w, cout), device=dev).to(torch.bfloat16))
        wshape = (cout, cin, 3, 3, 3) if is3d else (cout, cin, 3, 3)
        k = (3, 3, 3) if is3d else (1, 3, 3)
        flops = 2.0 * n * d * h * w * cin * cout * (27 if is3d else 9)
        res = {}
        for wide in (0, 1, 2, 3):
            _lib.lib().b200_set_tuning(6, wide)
            for _ in range(2):
                dw = ops.conv_bwd_weight(x, dy, k, wshape, impl=1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv_bwd_weight(x, dy, k, wshape, impl=1)
            e1.record()
            torch.cuda.synchronize()
            res[wide] = (e0.elapsed_time(e1) / 5, dw)
        _lib.lib().b200_set_tuning(6, 1)
        err = relerr(res[1][1], res[0][1])
        print(f"[PERF] wgrad {name}: narrow {res[0][0]:.3f} ms {flops / res[0][0] / 1e9:.0f} TF/s | "
              f"wide {res[1][0]:.3f} ms {flops / res[1][0] / 1e9:.0f} TF/s | pairs-only {res[2][0]:.3f} ms "
              f"centre-only {res[3][0]:.3f} ms | wide-vs-narrow rel {err[0]:.2e}",
              flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", required=True)
    a = ap.parse_args()
    t0 = time.time()
    print(f"== group {a.group} on {torch.cuda.get_device_name(0)}", flush=True)
    if a.group == "direct_fwd":
        ok = group_conv_fwd([2])
    elif a.group == "tc_fwd":
        ok = group_conv_fwd([1])
    elif a.group == "direct_bwd":
        ok = group_conv_bwd([2])
    elif a.group == "tc_bwd":
        ok = group_conv_bwd([1])
    elif a.group == "direct_convt":
        ok = group_convt([2])
    elif a.group == "tc_convt":
        ok = group_convt([1])
    elif a.group == "deconv":
        ok = group_deconv([2, 1])
    elif a.group == "perf":
        group_perf()
    elif a.group == "wgrad_perf":
        group_wgrad_perf()
        ok = True
    else:
        raise SystemExit("unknown group")
    print(f"== group {a.group} done ok={ok} in {time.time() - t0:.1f}s", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
