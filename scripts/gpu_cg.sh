#!/bin/bash
# A/B of the commit-group knobs (IMK_BT_CG1 / IMK_BT_CG3: blocks per tcgen05.commit of the 1x1 stages)
TAG=${1:-cg}; shift
CFGS=${@:-"isic2 hela"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
B="python bench.py --steps 5 --no-cpu-baseline --no-other-configs"
for c in $CFGS; do
  for v in "1 1" "2 2" "4 4" "8 8"; do
    set -- $v
    IMK_BT_CG1=$1 IMK_BT_CG3=$2 $B --config $c > $OUT/${c}_$1_$2.json 2>>$OUT/err.log
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${c}_$1_$2.json")); print("$c","cg1=$1 cg3=$2",round(d["value"]),[(k["kernel"],k["layer"],round(k["avg_us"])) for k in d["kernels"][:6]])
except Exception as e: print("ERR", e)
PY
  done
done
