#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/round4.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -40 >> gpurun_out/round4.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/round4.log
echo "=== layers" >> gpurun_out/round4.log
timeout 300 python tools/gpu_check.py layers 64 2000 >> gpurun_out/round4.log 2>&1
echo "=== bench ours" >> gpurun_out/round4.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "rc=$?" >> gpurun_out/round4.log
cat gpurun_out/bench4.json >> gpurun_out/round4.log; tail -5 gpurun_out/bench4.err >> gpurun_out/round4.log
tail -c 9000 gpurun_out/round4.log
