"""Top stall-sample SASS lines of an ncu report: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_src_top.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for idx, r in enumerate(rows[2:]):
    try:
        data.append((int(r[ci["# Samples"]]), idx, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, idx, r in sorted(data, key=lambda x: -x[0])[:N]:
    st = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{s:6d} {100.0*s/tot:5.1f}% #{idx:5d} exec={r[ci['Instructions Executed']]:>8} {r[ci['Source']][:70]:70s} {st}")
