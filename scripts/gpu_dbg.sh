#!/bin/bash
OUT=gpurun_out/${1:-dbg}
mkdir -p $OUT
export IMK_EXPECT_GPU=1
for k in 7; do
  IMK_BT_KINDS=$k timeout 40 python tools/fused_check.py 256 256 1 3 1.0 64 >> $OUT/dbg.log 2>&1; echo "hela64 kinds=$k exit $?" >> $OUT/dbg.log
done
grep -v "^$" $OUT/dbg.log | tail -30
if grep -q "exit 124" $OUT/dbg.log; then echo "HANG - stopping"; exit 1; fi
IMK_BT_TIMELINE=1 timeout 60 python tools/trunk_probe.py --config hela --images 64 --passes 2 --engine fused > $OUT/tl.log 2>&1; echo "tl exit $?"
timeout 300 python -m pytest tests/test_gpu_unet.py -x -q > $OUT/pytest_unet.log 2>&1; echo "pytest unet exit $?"; tail -3 $OUT/pytest_unet.log
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "im", d["roofline_im"]["frac"])
for k in d["kernels"]: print(k)
PY
tail -3 $OUT/bench.err
