#!/bin/bash
# same-box A/B: libimk_old.so (kernels of the previous commit) vs the current build, with and without the head stage
TAG=${1:-ab}; shift
CFGS=${@:-"hela isic5"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
B="python bench.py --steps 5 --no-cpu-baseline --no-other-configs"
for c in $CFGS; do
  [ -f inconsistencymasks_b200/libimk_old.so ] && IMK_LIB=$PWD/inconsistencymasks_b200/libimk_old.so $B --config $c > $OUT/${c}_old.json 2>>$OUT/err.log
  IMK_BT_NO_HEAD=1 $B --config $c > $OUT/${c}_nohead.json 2>>$OUT/err.log
  $B --config $c > $OUT/${c}_head.json 2>>$OUT/err.log
done
IMK_BT_VERBOSE=1 IMK_BT_TIMELINE=1 timeout 120 python tools/trunk_probe.py --config hela --images 64 --passes 2 --engine fused > $OUT/tl_hela.log 2>&1
python -m pytest tests/test_gpu_unet.py -x -q -m gpu 2>&1 | tail -2
python - <<PY
import json
for c in "$CFGS".split():
    for v in ("old","nohead","head"):
        try:
            d=json.load(open(f"$OUT/{c}_{v}.json"))
            print(c,v,round(d["value"]),round(d["e2e"]["value"]),[(k["kernel"],k["layer"],round(k["avg_us"])) for k in d["kernels"][:7]])
        except Exception as e: print(c,v,"ERR",e)
PY
