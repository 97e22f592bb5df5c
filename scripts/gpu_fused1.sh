#!/bin/bash
TAG=${1:-fused1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
IMK_BT_VERBOSE=1 timeout 300 python tools/trunk_probe.py --config hela --images 8 --passes 1 --engine fused > $OUT/probe.log 2>&1; echo "probe exit $?"; tail -12 $OUT/probe.log
timeout 600 python -m pytest tests/test_gpu_unet.py -x -q -s > $OUT/pytest_unet.log 2>&1; echo "pytest unet exit $?"; tail -15 $OUT/pytest_unet.log
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_all.log 2>&1; echo "pytest all exit $?"; tail -3 $OUT/pytest_all.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "im", d["roofline_im"]["frac"])
for k in d["kernels"]: print(k)
PY
tail -3 $OUT/bench.err
