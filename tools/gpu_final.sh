#!/bin/bash
# round-end checks on the GPU box: full -m gpu suite, default bench (with the CPU baseline), reference arm, smoke()
mkdir -p gpurun_out
LOG=gpurun_out/final.log
echo "=== pytest gpu" > $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -4 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
echo "=== bench default" >> $LOG
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?" >> $LOG
tail -c 1500 gpurun_out/bench_final.json >> $LOG
echo "=== bench reference arm" >> $LOG
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "rc=$?" >> $LOG
cat gpurun_out/bench_ref_final.json >> $LOG
cat $LOG
