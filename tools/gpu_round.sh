#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/round10.log
echo "=== pytest gpu" > $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -8 >> $LOG
echo "=== bench ours" >> $LOG
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "rc=$?" >> $LOG
python -c "import sys,json; d=json.loads(open('gpurun_out/bench10.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['stage_ms'])" >> $LOG 2>&1
tail -5 gpurun_out/bench10.err >> $LOG
cat $LOG
TAG=${TAG:-r01d} bash tools/gpu_launchlist.sh
