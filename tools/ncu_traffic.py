#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture of one trunk pass of the default bench shape:

    ncu --set full --clock-control none -k regex:block_tc -c 8 -o gpurun_out/X/traffic python tools/trunk_probe.py --config isic --images 512 --passes 1 --engine fused
    python tools/ncu_traffic.py gpurun_out/X/traffic.ncu-rep 512

DRAM bytes (read + write) per image of every block-fused launch, keyed like bench.py's kernel rows ("block_front:0", ...),
together with the digest of the sources the captured library was built from: bench.py reports `roofline.traffic` only when
that digest equals the digest of the library it runs (never a stale figure)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inconsistencymasks_b200 import build  # noqa: E402

rep, images = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# launch order of one fused trunk pass: FRONT, ENC 1..3 (blocks that are fused), DEC 3..0
keys = ["block_front:0", "block_enc:3", "block_enc:5", "block_enc:7", "block_dec:11", "block_dec:14", "block_dec:17", "block_dec:20"]
kernels = {}
for key, r in zip(keys, data):
    b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    kernels[key] = dict(bytes_per_image=b / images, kernel=r[ik][:80])
json.dump(dict(sources_digest=build.sources_digest(), capture=os.path.basename(rep), images=images, shape="ISIC 256x256x3, alpha 0.5 (default bench)",
               note="capture at the launch size of the bench (512 images); at 64 images the 126 MB L2 still holds part of a launch's output when it ends and the figures come out ~2x lower",
               kernels=kernels), open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(kernels, indent=1))
