#!/bin/bash
TAG=${1:-im}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
timeout 900 python -m pytest tests/test_gpu_im.py tests/test_gpu_unet.py -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest.log
for cfg in hela isic2 isic5 suim cityscapes; do timeout 120 python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
cat $OUT/im_bench.jsonl
