#!/bin/bash
TAG=${1:-im}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
timeout 300 python -m pytest tests/test_gpu_im.py -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest.log
for cfg in hela isic2 isic5 suim cityscapes; do timeout 60 python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
cat $OUT/im_bench.jsonl
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"])
for k in d["kernels"][:6]: print(k)
PY
