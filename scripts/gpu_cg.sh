#!/bin/bash
# A/B of the commit-group knobs (IMK_BT_CG1 / IMK_BT_CG3) + block-by-block issue timeline
TAG=${1:-cg}; OUT=gpurun_out/$TAG; mkdir -p $OUT
B="python bench.py --steps 5 --no-cpu-baseline --no-other-configs"
for c in hela isic5; do
  for v in "1 1" "4 4" "2 2" "4 1" "1 4" "8 8"; do
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
for v in "1 1" "4 4"; do set -- $v
IMK_BT_CG1=$1 IMK_BT_CG3=$2 IMK_LIB=$PWD/inconsistencymasks_b200/libimk_tl.so IMK_BT_VERBOSE=1 IMK_BT_TIMELINE=1 IMK_BT_TL_SKIP=8 timeout 120 python tools/trunk_probe.py --config hela --images 64 --passes 1 --engine fused > $OUT/tl_hela_$1_$2.log 2>&1
done
IMK_BT_CG1=4 IMK_BT_CG3=4 python -m pytest tests/test_gpu_unet.py -x -q -m gpu 2>&1 | tail -2
