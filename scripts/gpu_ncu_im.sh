#!/bin/bash
OUT=gpurun_out/${1:-ncuim}
mkdir -p $OUT
timeout 200 ncu --set full --clock-control none --import-source on -k regex:im_multiclass_tma -s 3 -c 1 -o $OUT/im_suim python tools/im_kernel_bench.py --config suim --images 256 --iters 5 > $OUT/ncu_suim.log 2>&1; tail -1 $OUT/ncu_suim.log
