#!/bin/bash
# Round-2 evidence run: GPU tests + smoke + bench (default + every --config, both arms) [+ ncu with NCU=1].
#   scripts/gpu_r2.sh TAG [configs...]
TAG=${1:-r2}; shift
CFGS=${@:-"isic5 hela isic2 suim city city2"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export IMK_EXPECT_GPU=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
rm -f gpurun_out/parity_table.jsonl
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
  cp gpurun_out/parity_table.jsonl $OUT/ 2>/dev/null
  timeout 120 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
fi
if [ -z "$SKIP_DEFAULT" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -2 $OUT/bench.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref exit $?"
fi
for cfg in $CFGS; do
  timeout 400 python bench.py --config $cfg --steps 5 --warmup 3 ${NO_CPU:+--no-cpu-baseline} > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "bench $cfg exit $?"
done
python - <<PY
import json, glob
for p in sorted(glob.glob("$OUT/bench*.json")):
    try:
        d = json.load(open(p))
    except Exception as e:
        print(p, "unreadable", e); continue
    if d.get("impl") == "reference":
        print(p, "reference", round(d["value"], 1), d["cpu_baseline"]["cores"], "cores"); continue
    r = d["roofline"]
    print(p, d["config"]["name"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "top", r["kernel"], r["layer"], r.get("frac"), "step", round(d["roofline_step"]["frac"], 3),
          "im", round(d["roofline_im"]["frac"], 3), "cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"], 1), round(d["cpu_baseline"]["batch64"]["value"], 1)))
    for k in d["kernels"][:8]: print("    ", k["kernel"], k["layer"], round(k["share"], 3), round(k["avg_us"]), k.get("gbs") and round(k["gbs"]))
    for o in d.get("configs", []): print("   other", o.get("name"), o.get("value") and round(o["value"]), o.get("e2e") and round(o["e2e"]), o.get("error"))
PY
if [ -n "$NCU" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/ncu_launches.csv python bench.py --steps 2 --warmup 1 --images-per-step 256 --e2e-images 64 --im-images 64 --no-cpu-baseline --no-other-configs > $OUT/ncu_launches.log 2>&1; echo "ncu launches exit $?"
  for cfg in ${NCU_CFGS:-isic hela}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:block_tc -s 6 -c 6 -o $OUT/block_tc_$cfg python tools/trunk_probe.py --config $cfg --engine fused > $OUT/ncu_block_$cfg.log 2>&1; echo "ncu block $cfg exit $?"
  done
fi
ls -la $OUT
