#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/round12.log
: > $LOG
for m in 0 auto 1; do
  echo "=== DCS_EARLY_SKIP=$m" >> $LOG
  DCS_EARLY_SKIP=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])" >> $LOG 2>&1
done
echo "=== pytest gpu" >> $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -4 >> $LOG
cat $LOG
