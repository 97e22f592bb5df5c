#!/bin/bash
# tests + smoke + bench + IM kernel bench (every step under its own timeout)
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
for cfg in hela isic2 isic5 suim cityscapes; do timeout 60 python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
cat $OUT/im_bench.jsonl
timeout 300 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "im", d["roofline_im"]["frac"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
for k in d["kernels"]: print(k)
PY
tail -3 $OUT/bench.err
