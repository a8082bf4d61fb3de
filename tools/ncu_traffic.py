"""DRAM bytes per C-ABI call of the two tensor-kernel families from an ncu launch list (see tools/ncu_summary.py for
the csv): python tools/ncu_traffic.py profiles/r02_cfg3_launches_final.csv cfg3 26 14  ->  JSON fragment.
The number of steps in the list is the number of loss-sum launches (one per step)."""
import collections
import csv
import json
import sys

path, workload, gemm_calls, wgrad_calls = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
byt = collections.defaultdict(float)
cnt = collections.Counter()
seen = set()
for r in data:
    if len(r) <= vi:
        continue
    name = r[ki]
    fam = ("tc_wgrad_kernel" if "tc_wgrad" in name else "tc_gemm_kernel" if "tc_gemm" in name
           else "step" if ("seg_sums" in name or "ce_fwd" in name) else None)
    if fam is None:
        continue
    if (r[idi], fam) not in seen:
        seen.add((r[idi], fam))
        cnt[fam] += 1
    if r[mi].startswith("dram__bytes"):
        byt[fam] += float(r[vi].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
steps = cnt["step"]
out = {workload: {
    "steps_in_list": steps,
    "tc_gemm_kernel": {"kernel_launches": cnt["tc_gemm_kernel"], "calls_per_step": gemm_calls,
                       "dram_bytes_per_launch": byt["tc_gemm_kernel"] / max(1, steps * gemm_calls)},
    "tc_wgrad_kernel": {"kernel_launches": cnt["tc_wgrad_kernel"], "calls_per_step": wgrad_calls,
                        "dram_bytes_per_launch": byt["tc_wgrad_kernel"] / max(1, steps * wgrad_calls)}}}
print(json.dumps(out, indent=1))
