#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/round3.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -80 >> gpurun_out/round3.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/round3.log
echo "=== bf16 taps (synthetic weights)" >> gpurun_out/round3.log
timeout 300 python tools/gpu_check.py bf16 2 256 >> gpurun_out/round3.log 2>&1
echo "=== bench ours" >> gpurun_out/round3.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "rc=$?" >> gpurun_out/round3.log
cat gpurun_out/bench3.json >> gpurun_out/round3.log; tail -5 gpurun_out/bench3.err >> gpurun_out/round3.log
echo "=== bench fp32" >> gpurun_out/round3.log
timeout 600 python bench.py --steps 5 --warmup 3 --mode fp32 --no-cpu-baseline > gpurun_out/bench3_fp32.json 2> gpurun_out/bench3_fp32.err; echo "rc=$?" >> gpurun_out/round3.log
cat gpurun_out/bench3_fp32.json >> gpurun_out/round3.log; tail -5 gpurun_out/bench3_fp32.err >> gpurun_out/round3.log
tail -c 9000 gpurun_out/round3.log
