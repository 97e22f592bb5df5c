#!/bin/bash
OUT=gpurun_out/${1:-tlo}
mkdir -p $OUT
IMK_BT_TIMELINE=1 timeout 60 python tools/trunk_probe.py --config hela --images 64 --passes 2 --engine fused > $OUT/tl.log 2>&1; echo "tl exit $?"
grep -A26 "timeline kind=2 256x256" $OUT/tl.log | tail -27 | sed -n '19,24p'
