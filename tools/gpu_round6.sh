#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/round6.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -30 >> gpurun_out/round6.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/round6.log
echo "=== layers" >> gpurun_out/round6.log
timeout 300 python tools/gpu_check.py layers 64 2000 >> gpurun_out/round6.log 2>&1
echo "=== bench ours" >> gpurun_out/round6.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "rc=$?" >> gpurun_out/round6.log
cat gpurun_out/bench6.json >> gpurun_out/round6.log; tail -5 gpurun_out/bench6.err >> gpurun_out/round6.log
echo "=== ncu launch list (eager pass)" >> gpurun_out/round6.log
timeout 300 python tools/prof_pass.py > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 69 -c 80 --csv --log-file gpurun_out/launches.csv python tools/prof_pass.py > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/round6.log
tail -c 7000 gpurun_out/round6.log
