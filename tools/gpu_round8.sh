#!/bin/bash
# strip kernel bring-up: small unit check, gpu tests, full-size unit check + timing, bench
mkdir -p gpurun_out
LOG=gpurun_out/round8.log
echo "=== stripunit small" > $LOG
timeout 120 python tools/gpu_check.py stripunit 2 256 >> $LOG 2>&1; echo "rc=$?" >> $LOG
echo "=== pytest gpu" >> $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -15 >> $LOG
echo "=== stripunit full" >> $LOG
timeout 300 python tools/gpu_check.py stripunit 64 2000 >> $LOG 2>&1; echo "rc=$?" >> $LOG
echo "=== bench ours" >> $LOG
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "rc=$?" >> $LOG
cat gpurun_out/bench8.json >> $LOG; tail -5 gpurun_out/bench8.err >> $LOG
tail -c 7000 $LOG
