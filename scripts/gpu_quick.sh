#!/bin/bash
# quick iteration: GPU parity tests + IM kernel bench + bench (no CPU baseline) + block-fused timeline
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
timeout 300 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
for cfg in hela isic2 isic5 suim cityscapes; do timeout 60 python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
python - <<PY
import json
for l in open("$OUT/im_bench.jsonl"):
    d=json.loads(l); print(d["config"], round(d["gbs"]), round(d["frac"],3))
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "im", d["roofline_im"]["frac"])
for k in d["kernels"]: print(k)
PY
tail -3 $OUT/bench.err
IMK_TC_VERBOSE=1 IMK_BT_VERBOSE=1 IMK_BT_TIMELINE=1 timeout 60 python tools/trunk_probe.py --config hela --images 64 --passes 2 --engine fused > $OUT/tl.log 2>&1; echo "tl exit $?"
