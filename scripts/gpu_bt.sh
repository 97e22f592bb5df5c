#!/bin/bash
# block-fused engine: timeline of CTA 0 + ncu full capture of the 6 block_tc launches of one HeLa trunk pass
OUT=gpurun_out/${1:-bt}
mkdir -p $OUT
IMK_BT_VERBOSE=1 IMK_BT_TIMELINE=1 timeout 60 python tools/trunk_probe.py --config hela --images 64 --passes 2 --engine fused > $OUT/tl.log 2>&1; echo "tl exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_tc -s 6 -c 6 -o $OUT/block_tc_hela python tools/trunk_probe.py --config hela --engine fused > $OUT/ncu_block.log 2>&1; echo "ncu block exit $?"; tail -2 $OUT/ncu_block.log
ls -la $OUT
