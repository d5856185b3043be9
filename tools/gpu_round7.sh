#!/bin/bash
# round-1 evidence pass: descriptor probe, gpu tests, bench, ncu launch list + full-section capture of one eager pass
mkdir -p gpurun_out
LOG=gpurun_out/round7.log
echo "=== umma shift probe" > $LOG
timeout 120 ./tools/umma_shift_test > gpurun_out/umma_shift.log 2>&1; echo "rc=$?" >> $LOG
tail -3 gpurun_out/umma_shift.log >> $LOG
echo "=== pytest gpu" >> $LOG
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -5 >> $LOG
echo "=== bench ours" >> $LOG
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "rc=$?" >> $LOG
cat gpurun_out/bench7.json >> $LOG; tail -5 gpurun_out/bench7.err >> $LOG
echo "=== ncu" >> $LOG
timeout 300 python tools/prof_pass.py > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 71 -c 71 --csv --log-file gpurun_out/launches7.csv python tools/prof_pass.py > gpurun_out/ncu_launches.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -s 71 -c 71 -o gpurun_out/prof_r01_full -f python tools/prof_pass.py > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?" >> $LOG
ls -la gpurun_out/*.ncu-rep >> $LOG
tail -c 6000 $LOG
