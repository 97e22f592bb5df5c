#!/bin/bash
TAG=${1:-ncu2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_tc -s 6 -c 6 -o $OUT/block_tc_hela python tools/trunk_probe.py --config hela --engine fused > $OUT/ncu_block.log 2>&1; tail -2 $OUT/ncu_block.log
ls -la $OUT
