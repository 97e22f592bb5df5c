#!/bin/bash
# Final evidence run of a round: scripts/gpu_r2.sh (tests, smoke, both bench arms, every config) + the ncu launch list of the
# bench command + one `--set full` capture of a fused trunk pass (DRAM traffic per launch -> profiles/ncu_traffic.json).
TAG=${1:-final}
bash scripts/gpu_r2.sh $TAG
OUT=gpurun_out/$TAG
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/ncu_launches.csv python bench.py --steps 2 --warmup 1 --images-per-step 256 --e2e-images 64 --im-images 64 --no-cpu-baseline --no-other-configs > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_tc -c 8 -o $OUT/traffic python tools/trunk_probe.py --config isic --images 512 --passes 1 --engine fused > $OUT/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
python tools/ncu_traffic.py $OUT/traffic.ncu-rep 512 > $OUT/traffic.json 2>$OUT/traffic.err; cp profiles/ncu_traffic.json $OUT/ncu_traffic.json
ls -la $OUT
