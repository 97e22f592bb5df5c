#!/bin/bash
# knock-out experiment (timeline build): which role bounds the block kernels?  results of the runs are wrong by design
TAG=${1:-ko}; shift
CFGS=${@:-"isic2 hela"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
for c in $CFGS; do
  for k in 0 1 2 4 8 24 32 6 63; do
    IMK_BT_KNOCK=$k IMK_LIB=$PWD/inconsistencymasks_b200/libimk_tl.so timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-other-configs --config $c > $OUT/${c}_$k.json 2>>$OUT/err.log
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${c}_$k.json")); print("$c knock=$k",round(d["value"]),[(k["kernel"],k["layer"],round(k["avg_us"])) for k in d["kernels"][:7]])
except Exception as e: print("ERR", e)
PY
  done
done
