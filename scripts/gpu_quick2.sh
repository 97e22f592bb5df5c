#!/bin/bash
# quick check: unet + parity tests, then per-kernel bench lines for the given configs
TAG=${1:-q}; shift
CFGS=${@:-"isic5 hela"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_parity_baseline.py -x -q -m gpu 2>&1 | tail -${TAIL:-4}; fi
for c in $CFGS; do
  python bench.py --steps 5 --no-cpu-baseline --no-other-configs --config $c > $OUT/$c.json 2>>$OUT/err.log
  python - <<PY
import json
try:
    d=json.load(open("$OUT/$c.json")); print("$c",round(d["value"]),round(d["e2e"]["value"]),[(k["kernel"],k["layer"],round(k["avg_us"])) for k in d["kernels"][:9]])
except Exception as e: print("ERR", e)
PY
done
tail -3 $OUT/err.log
