#!/bin/bash
TAG=${1:-ncu1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python tools/pcie_probe.py > $OUT/pcie.json 2>&1; cat $OUT/pcie.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 22 -c 22 -o $OUT/conv_tc_hela python tools/trunk_probe.py --config hela > $OUT/ncu_conv.log 2>&1; tail -2 $OUT/ncu_conv.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:in_conv -s 1 -c 1 -o $OUT/in_conv_hela python tools/trunk_probe.py --config hela > $OUT/ncu_inconv.log 2>&1; tail -2 $OUT/ncu_inconv.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:im_multiclass_tma -s 3 -c 1 -o $OUT/im_suim python tools/im_kernel_bench.py --config suim --images 256 --iters 5 > $OUT/ncu_suim.log 2>&1; tail -2 $OUT/ncu_suim.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:im_binary_vec -s 3 -c 1 -o $OUT/im_hela python tools/im_kernel_bench.py --config hela --images 512 --iters 5 > $OUT/ncu_hela.log 2>&1; tail -2 $OUT/ncu_hela.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ensemble_im -s 1 -c 1 -o $OUT/ens_hela python bench.py --steps 1 --warmup 1 --images-per-step 64 --e2e-images 64 --im-images 64 --no-cpu-baseline > $OUT/ncu_ens.log 2>&1; tail -2 $OUT/ncu_ens.log
ls -la $OUT
