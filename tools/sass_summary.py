#!/usr/bin/env python
"""profiles/sass_summary.txt: opcode histogram of libimk.so per kernel (cuobjdump -sass; runs without a GPU).

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("IMK_LIB", os.path.join(ROOT, "inconsistencymasks_b200", "libimk.so"))
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
filt = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout or out
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCATOMSWS", "SYNCS", "STS", "LDS", "LDG", "STG", "ATOMS", "HMMA", "LDSM", "MUFU"]
rows, name, hist, n = [], None, None, 0
def flush():
    if name is not None:
        rows.append((n, " ".join(f"{k}={hist[k]}" for k in KEYS if hist[k]), name))
for line in filt.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        flush(); name, hist, n = m.group(1).strip(), collections.Counter(), 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name is not None:
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k):
                hist[k] += 1; break
flush()
print("# SASS opcode histogram of libimk.so (cuobjdump -sass), sm_100a; per kernel: instructions | tcgen05 / TMA / mbarrier opcodes")
for n, h, name in sorted(rows, key=lambda r: -r[0]):
    print(f"{n:7d}  {h}  | {name}")
