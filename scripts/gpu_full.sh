#!/bin/bash
# Round evidence run: tests + smoke + IM kernel bench + bench (both arms) + ncu launch list + ncu full of the top kernels.
TAG=${1:-full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
for cfg in hela isic2 isic5 suim cityscapes; do timeout 60 python tools/im_kernel_bench.py --config $cfg --images 512 >> $OUT/im_bench.jsonl 2>> $OUT/im_bench.err; done
cat $OUT/im_bench.jsonl
timeout 400 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "im", d["roofline_im"]["frac"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
for k in d["kernels"]: print(k)
PY
tail -3 $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"; cat $OUT/bench_reference.json
# the other BASELINE configs through the fused path (parity-test shapes; throughput for DESIGN.md)
for cfg in hela isic isic5 suim city city2; do timeout 120 python tools/ensemble_bench.py --config $cfg --images 512 >> $OUT/configs.jsonl 2>> $OUT/configs.err; done
python - <<PY
import json
for l in open("$OUT/configs.jsonl"):
    d=json.loads(l); print(d["config"], round(d["images_per_s"]), [ (k["kernel"], k["layer"], k["share"]) for k in d["kernels"][:3]])
PY
# ncu launch list of the bench command (share of step per kernel)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/ncu_launches.csv python bench.py --steps 2 --warmup 1 --images-per-step 256 --e2e-images 64 --im-images 64 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
# ncu full: block-fused trunk kernels (one pass of one model) and the fused epilogue / IM kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_tc -s 6 -c 6 -o $OUT/block_tc_hela python tools/trunk_probe.py --config hela --engine fused > $OUT/ncu_block.log 2>&1; echo "ncu block exit $?"; tail -2 $OUT/ncu_block.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ensemble_im -s 1 -c 1 -o $OUT/ens_hela python bench.py --steps 1 --warmup 1 --images-per-step 64 --e2e-images 64 --im-images 64 --no-cpu-baseline > $OUT/ncu_ens.log 2>&1; echo "ncu ens exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:im_binary_vec -s 3 -c 1 -o $OUT/im_hela python tools/im_kernel_bench.py --config hela --images 512 --iters 5 > $OUT/ncu_hela.log 2>&1; echo "ncu im hela exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:im_multiclass -s 3 -c 1 -o $OUT/im_suim python tools/im_kernel_bench.py --config suim --images 256 --iters 5 > $OUT/ncu_suim.log 2>&1; echo "ncu im suim exit $?"
ls -la $OUT
