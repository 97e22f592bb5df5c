#!/bin/bash
# bisect helper: bench one config with several prebuilt libimk variants (IMK_LIB)
TAG=$1; CFG=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  IMK_LIB=$PWD/inconsistencymasks_b200/libimk_$v.so python bench.py --steps 5 --no-cpu-baseline --no-other-configs --config $CFG > $OUT/${CFG}_$v.json 2>>$OUT/err.log
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${CFG}_$v.json")); print("$CFG","$v",round(d["value"]),[(k["kernel"],k["layer"],round(k["avg_us"])) for k in d["kernels"][:8]])
except Exception as e: print("$v ERR", e)
PY
done
