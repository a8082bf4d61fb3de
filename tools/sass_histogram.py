"""SASS opcode histogram of libb200unet.so (cuobjdump -sass): the Blackwell-native evidence the profiling recipe
asks for (UTCHMMA = tcgen05.mma, UTMALDG = TMA tile loads, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, ...).
python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mdeical_image_segmentation_b200", "libb200unet.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
ops = collections.Counter()
per_kernel = collections.defaultdict(collections.Counter)
cur = "?"
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        op = m.group(1)
        ops[op.split(".")[0]] += 1
        per_kernel[cur][op.split(".")[0]] += 1
KEY = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA",
       "HGMMA", "RED", "REDG", "ATOMG", "LDG", "STG", "LDS", "STS", "ELECT"]
print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {sum(ops.values())} instructions in {len(per_kernel)} kernels")
print("\n## Blackwell-native opcodes (whole library)")
for k in KEY:
    print(f"{k:12s} {ops.get(k, 0):8d}")
print("\n## per kernel: UTCHMMA / UTMALDG / LDTM / UTCBAR (kernels with any)")
for name, c in sorted(per_kernel.items(), key=lambda kv: -kv[1].get("UTCHMMA", 0)):
    if c.get("UTCHMMA", 0) or c.get("UTMALDG", 0):
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:90]
        print(f"{c.get('UTCHMMA', 0):6d} {c.get('UTMALDG', 0):5d} {c.get('LDTM', 0):5d} {c.get('UTCBAR', 0):5d}  {short}")
print("\n## top 25 opcodes")
for op, n in ops.most_common(25):
    print(f"{op:12s} {n:8d}")
