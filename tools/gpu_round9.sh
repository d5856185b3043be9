#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/round9.log
echo "=== pytest gpu" > $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -15 >> $LOG
echo "=== layers" >> $LOG
timeout 300 python tools/gpu_check.py layers 64 2000 2>&1 | grep -E "\"op\"|rror" >> $LOG
echo "=== bench ours" >> $LOG
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "rc=$?" >> $LOG
cat gpurun_out/bench9.json >> $LOG; tail -5 gpurun_out/bench9.err >> $LOG
echo "=== bench unfused attention" >> $LOG
DCS_FUSED_ATTENTION=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['stage_ms'])" >> $LOG
tail -c 7000 $LOG
