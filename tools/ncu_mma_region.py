import csv, io, subprocess, sys
rep=sys.argv[1]; idx=int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))   # first (SASS) section only
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:end]))))
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"] or 0) for r in rows)
# find MMA instructions
mma=[i for i,r in enumerate(rows) if "UTCHMMA" in r["Source"] or "UTCBAR" in r["Source"]]
print("rows",len(rows),"tot",tot,"mma range",mma[0],mma[-1])
lo,hi=mma[0]-150,mma[-1]+60
agg={c:0 for c in stall_cols}; ssum=0
for r in rows[lo:hi]:
    s=int(r["# Samples"] or 0); ssum+=s
    for c in stall_cols: agg[c]+=int(r[c] or 0)
print("samples in MMA-issue code region:", ssum, ssum/tot)
print({k[6:]:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:8]})
for i in range(lo,hi):
    r=rows[i]; s=int(r["# Samples"] or 0)
    if s>=25:
        main=max(stall_cols,key=lambda c:int(r[c] or 0))
        print(i,s,main[6:],r["Instructions Executed"],r["Source"].strip()[:90])
