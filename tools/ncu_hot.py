#!/usr/bin/env python
"""Top stalled SASS instructions of one launch: python tools/ncu_hot.py rep.ncu-rep [launch_idx] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))   # first (SASS) section only
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
print("stall mix:", ", ".join(f"{k[6:]}={v/tot:.2f}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for i, r in enumerate(rows): r["_i"] = i
for r in sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:top]:
    s = int(r["# Samples"] or 0)
    main = max(stall_cols, key=lambda c: int(r[c] or 0))
    print(f"{r['_i']:5d} {s/tot:6.3f} {main[6:]:14s} exec={r['Instructions Executed']:>9s} {r['Source'].strip()[:100]}")
