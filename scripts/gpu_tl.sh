#!/bin/bash
TAG=${1:-tl}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
IMK_BT_TIMELINE=1 timeout 300 python tools/trunk_probe.py --config hela --images 64 --passes 2 --engine fused > $OUT/tl.log 2>&1; echo "tl exit $?"
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q > $OUT/pytest_unet.log 2>&1; echo "pytest unet exit $?"; tail -3 $OUT/pytest_unet.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "im", d["roofline_im"]["frac"])
for k in d["kernels"]: print(k)
PY
tail -3 $OUT/bench.err
