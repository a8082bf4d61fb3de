"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list per kernel:
launch count, total device time, share of the listed time and DRAM bytes (cold-cache, serialised)."""
import collections
import csv
import sys


def main(path, skip=0):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, mi, vi, ui, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    names = {}
    for r in data:
        if len(r) <= vi:
            continue
        lid = int(r[idi])
        if lid < skip:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        if r[mi].startswith("gpu__time"):
            v = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
        else:
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        per[lid][r[mi]] = v
        names[lid] = r[ki].split("(")[0].replace("void ", "")[:48]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for lid, m in per.items():
        a = agg[names[lid]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(per)} launches, {tot:.2f} ms of kernel time listed")
    print(f"{'kernel':50s} {'n':>5s} {'ms':>9s} {'share':>6s} {'dram GB':>8s} {'GB/s':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
        gbs = v[2] / v[1] / 1e6 if v[1] else 0
        print(f"{k:50s} {v[0]:5d} {v[1]:9.3f} {v[1] / tot:6.3f} {v[2] / 1e9:8.2f} {gbs:7.0f}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
