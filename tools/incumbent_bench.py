#!/usr/bin/env python
"""GPU incumbent (SURVEY.md §8d "the kernel to beat"): the same architectures built from STOCK PyTorch
modules (nn.Conv*/GroupNorm/MaxPool/ConvTranspose -> ATen -> cuDNN 9 / cuBLAS), channels-last, bf16
autocast, cuDNN benchmark on, timed with CUDA events on the same box as bench.py.  Self-contained on
purpose: no oracle/, no product code — only torch.

    python tools/incumbent_bench.py --workload cfg2|cfg3|cfg3p|cfg4 [--steps 10 --warmup 5]
"""
from __future__ import annotations

import argparse
import json

import torch
import torch.nn as nn
import torch.nn.functional as F


class Stock2D(nn.Module):
    """model/unet2d/unet.py:42-128 with stock modules (conv-ReLU x2, pool, ConvT k2s2, cat([up, skip]))."""

    def __init__(self, cin, cout):
        super().__init__()
        dc = lambda i, o: nn.Sequential(nn.Conv2d(i, o, 3, padding=1), nn.ReLU(),   # noqa: E731
                                        nn.Conv2d(o, o, 3, padding=1), nn.ReLU())
        chans = [(cin, 64), (64, 128), (128, 256), (256, 512)]
        self.down = nn.ModuleList([dc(i, o) for i, o in chans])
        self.mid = dc(512, 1024)
        ups = [(1024, 512), (512, 256), (256, 128), (128, 64)]
        self.up = nn.ModuleList([nn.ConvTranspose2d(i, o, 2, stride=2) for i, o in ups])
        self.upc = nn.ModuleList([dc(i, o) for i, o in ups])
        self.head = nn.Conv2d(64, cout, 1)

    def forward(self, x):
        skips = []
        for d in self.down:
            x = d(x)
            skips.append(x)
            x = F.max_pool2d(x, 2)
        x = self.mid(x)
        for u, c in zip(self.up, self.upc):
            x = torch.cat([u(x), skips.pop()], 1)
            x = c(x)
        return self.head(x)


class Stock3D(nn.Module):
    """model/unet3d/model.py:13-194: f_maps 64..512, 'gcr' (GroupNorm-conv-ReLU) + nearest, or
    'cbr' (conv-BatchNorm-ReLU) + ConvTranspose3d(k3,s2,p1) then nearest resize."""

    def __init__(self, cin, cout, order="gcr", upsample="default"):
        super().__init__()
        f = [64, 128, 256, 512]

        def single(i, o):
            if order == "gcr":
                return nn.Sequential(nn.GroupNorm(8 if i >= 8 else 1, i), nn.Conv3d(i, o, 3, padding=1, bias=False),
                                     nn.ReLU(inplace=True))
            return nn.Sequential(nn.Conv3d(i, o, 3, padding=1, bias=False), nn.BatchNorm3d(o), nn.ReLU(inplace=True))

        def enc(i, o):
            m = max(o // 2, i)
            return nn.Sequential(single(i, m), single(m, o))

        def dec(i, o):
            return nn.Sequential(single(i, o), single(o, o))

        self.enc = nn.ModuleList([enc(cin, f[0])] + [enc(f[k], f[k + 1]) for k in range(3)])
        self.deconv = upsample == "deconv"
        if self.deconv:
            self.ct = nn.ModuleList([nn.ConvTranspose3d(f[k + 1], f[k], 3, stride=2, padding=1, bias=False)
                                     for k in (2, 1, 0)])
            self.dec = nn.ModuleList([dec(2 * f[k], f[k]) for k in (2, 1, 0)])
        else:
            self.dec = nn.ModuleList([dec(f[k] + f[k + 1], f[k]) for k in (2, 1, 0)])
        self.head = nn.Conv3d(f[0], cout, 1)

    def forward(self, x):
        feats = []
        for k, e in enumerate(self.enc):
            if k:
                x = F.max_pool3d(x, 2)
            x = e(x)
            feats.append(x)
        feats.pop()
        for j, d in enumerate(self.dec):
            skip = feats.pop()
            if self.deconv:
                x = self.ct[j](x)
            x = F.interpolate(x, size=skip.shape[2:], mode="nearest")
            x = d(torch.cat((skip, x), 1))
        return self.head(x)


def bce_dice(logits, target):
    """model/unet3d/losses.py:167-178 (alpha = beta = 1)."""
    bce = F.binary_cross_entropy_with_logits(logits, target)
    p = torch.sigmoid(logits)
    c = logits.shape[1]
    pf = p.transpose(0, 1).reshape(c, -1)
    tf = target.transpose(0, 1).reshape(c, -1)
    dice = 2 * (pf * tf).sum(-1) / ((pf * pf).sum(-1) + (tf * tf).sum(-1)).clamp(min=1e-6)
    return bce + (1 - dice.mean())


def dice_only(logits, target):
    p = torch.sigmoid(logits)
    c = logits.shape[1]
    pf = p.transpose(0, 1).reshape(c, -1)
    tf = target.transpose(0, 1).reshape(c, -1)
    dice = 2 * (pf * tf).sum(-1) / ((pf * pf).sum(-1) + (tf * tf).sum(-1)).clamp(min=1e-6)
    return 1 - dice.mean()


CFG = {
    "cfg2": dict(kind="2d", cin=3, cout=4, size=(512, 512), batch=64, unit="imgs/s"),
    "cfg3": dict(kind="3d", cin=1, cout=1, size=(128,) * 3, batch=2, order="gcr", upsample="default", unit="vols/s"),
    "cfg3p": dict(kind="3d", cin=1, cout=1, size=(128,) * 3, batch=2, order="cbr", upsample="deconv", unit="vols/s"),
    "cfg4": dict(kind="3d", cin=4, cout=3, size=(160, 192, 160), batch=1, order="gcr", upsample="default",
                 unit="vols/s", loss="dice"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=sorted(CFG))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    w = CFG[args.workload]
    b = args.batch or w["batch"]
    torch.backends.cudnn.benchmark = True
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    if w["kind"] == "2d":
        net = Stock2D(w["cin"], w["cout"]).to(dev).to(memory_format=torch.channels_last)
        x = torch.randn((b, w["cin"]) + w["size"], device=dev).contiguous(memory_format=torch.channels_last)
        y = torch.randint(0, w["cout"], (b,) + w["size"], device=dev)
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-3, fused=True)
        loss_fn = lambda o: F.cross_entropy(o.float(), y)   # noqa: E731
    else:
        net = Stock3D(w["cin"], w["cout"], w["order"], w["upsample"]).to(dev).to(memory_format=torch.channels_last_3d)
        x = torch.randn((b, w["cin"]) + w["size"], device=dev).contiguous(memory_format=torch.channels_last_3d)
        y = (torch.rand((b, w["cout"]) + w["size"], device=dev) > 0.7).float()
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, fused=True)
        lf = dice_only if w.get("loss") == "dice" else bce_dice
        loss_fn = lambda o: lf(o.float(), y)   # noqa: E731
    net.train()

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net(x)
        loss = loss_fn(out)
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "incumbent: stock PyTorch modules, cuDNN, bf16 autocast, channels_last",
                      "workload": args.workload, "value": round(b / (ms / 1e3), 3), "unit": w["unit"],
                      "ms_per_step": round(ms, 3), "steps": args.steps, "warmup": args.warmup, "batch": b,
                      "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
                      "final_loss": float(loss.item()),
                      "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}), flush=True)


if __name__ == "__main__":
    main()
