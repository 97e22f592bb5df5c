#!/bin/bash
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
for cfg in hela suim cityscapes; do timeout 120 python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
cat $OUT/im_bench.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "im", d["roofline_im"]["frac"])
for k in d["kernels"]: print(k)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 130 -c 2 -o $OUT/conv_tc python bench.py --steps 1 --warmup 1 --images-per-step 128 --e2e-images 64 --im-images 64 --no-cpu-baseline > $OUT/ncu_conv.log 2>&1
tail -2 $OUT/ncu_conv.log
