#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both engines), ncu launch list + one full capture.
TAG=${1:-r01b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
timeout 1200 python -m pytest tests -x -q -m gpu -s > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log
tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
head -c 5000 $OUT/bench.json
tail -5 $OUT/bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --engine direct --no-cpu-baseline > $OUT/bench_direct.json 2> $OUT/bench_direct.err; echo "bench direct exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref exit $?"; cat $OUT/bench_ref.json
for cfg in hela isic2 isic5 suim cityscapes; do timeout 120 python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
cat $OUT/im_bench.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 3 --images-per-step 128 --e2e-images 64 --im-images 64 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "launch list exit $?"; wc -l $OUT/launches.csv
