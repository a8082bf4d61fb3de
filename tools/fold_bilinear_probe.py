"""CPU probe for DESIGN.md section 8 item 2: conv3x3(pad 1) o bilinear(x s, align_corners=False) equals, for every
output phase, ONE 3x3 filter on the replicate-padded low-resolution map; the filter depends on the phase
(s interior phases per axis + one variant for the first / last output row or column, where the convolution's zero
padding removes a tap).  Builds the (s+2)^2 effective filters from the original weights and checks the folded
evaluation against torch on random data.  Pure torch, runs anywhere:  python tools/fold_bilinear_probe.py"""
import torch
import torch.nn.functional as F


def axis_tables(s: int):
    """R[v][d][a]: weight of low-resolution index i + a - 1 (a = 0..2) in up-sampled index s*i + p + d - 1
    (d = 0..2 = conv tap), for the phase variants v = 0..s+1: v = 1..s interior phase p = v - 1; v = 0: first
    output index (p = 0, tap d = 0 falls into the zero padding); v = s + 1: last output index (p = s - 1, tap 2 out)."""
    R = torch.zeros(s + 2, 3, 3, dtype=torch.float64)
    for v in range(s + 2):
        p = 0 if v == 0 else (s - 1 if v == s + 1 else v - 1)
        for d in range(3):
            if (v == 0 and d == 0) or (v == s + 1 and d == 2):
                continue                                     # zero padding of the convolution
            t = p + d - 1                                    # offset of the up-sampled index from s*i
            src = (t + 0.5) / s - 0.5                        # relative to low-resolution index i
            i0 = int(torch.floor(torch.tensor(src)).item())
            l1 = src - i0
            R[v, d, i0 + 1] += 1.0 - l1
            R[v, d, i0 + 2] += l1
    return R


def folded(x, w, bias, s):
    n, c, h, wd = x.shape
    R = axis_tables(s).to(x.dtype)
    xp = F.pad(x, (1, 1, 1, 1), mode="replicate")           # edge clamping of the bilinear resize
    out = torch.empty((n, w.shape[0], h * s, wd * s), dtype=x.dtype)
    count = set()
    for Y in range(h * s):
        vy = 0 if Y == 0 else (s + 1 if Y == h * s - 1 else Y % s + 1)
        for X in range(wd * s):
            vx = 0 if X == 0 else (s + 1 if X == wd * s - 1 else X % s + 1)
            count.add((vy, vx))
            E = torch.einsum("ocde,da,eb->ocab", w, R[vy], R[vx])       # effective 3x3 filter of this phase
            i, j = Y // s, X // s
            out[:, :, Y, X] = torch.einsum("ncab,ocab->no", xp[:, :, i:i + 3, j:j + 3], E) + bias
    return out, len(count)


def main():
    torch.manual_seed(0)
    for s in (2, 4, 8):
        x = torch.randn(2, 5, 4, 3, dtype=torch.float64)
        w = torch.randn(6, 5, 3, 3, dtype=torch.float64)
        b = torch.randn(6, dtype=torch.float64)
        ref = F.conv2d(F.interpolate(x, scale_factor=s, mode="bilinear"), w, b, padding=1)
        got, nfilters = folded(x, w, b, s)
        err = (got - ref).abs().max().item()
        print(f"scale {s}: max |folded - conv(bilinear)| = {err:.2e} with {nfilters} distinct phase filters "
              f"(bound (s+2)^2 = {(s + 2) ** 2})")
        assert err < 1e-10


if __name__ == "__main__":
    main()
