#!/bin/bash
TAG=${1:-tc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
for v in 0 1; do
  IMK_TC_SWAP_LBO_SBO=$v timeout 300 python -m pytest tests/test_gpu_unet.py -x -q -s -k "predict_vs_fp32_oracle and tcgen05" > $OUT/tc_variant$v.log 2>&1
  echo "variant $v exit $?"; grep -E "max\|dp|passed|failed|Error|error" $OUT/tc_variant$v.log | head -12
done
