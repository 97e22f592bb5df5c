#!/bin/bash
TAG=${1:-im}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in hela isic2 isic5 suim cityscapes; do python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
cat $OUT/im_bench.jsonl
ncu --set full --clock-control none --import-source on -k regex:im_binary_vec -s 3 -c 1 -o $OUT/im_hela python tools/im_kernel_bench.py --config hela --images 512 --iters 5 > $OUT/ncu_hela.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:im_multiclass_tma -s 3 -c 1 -o $OUT/im_city python tools/im_kernel_bench.py --config cityscapes --images 256 --iters 5 > $OUT/ncu_city.log 2>&1
tail -3 $OUT/ncu_hela.log
