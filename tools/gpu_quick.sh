#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -5 > gpurun_out/quick.log
timeout 300 python tools/gpu_check.py layers 64 2000 2>&1 | grep -E "\"op\"|rror" | grep -v att >> gpurun_out/quick.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline >> gpurun_out/quick.log 2>&1
cat gpurun_out/quick.log
