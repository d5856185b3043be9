#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/round2.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -60 >> gpurun_out/round2.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/round2.log
echo "=== layers" >> gpurun_out/round2.log
timeout 300 python tools/gpu_check.py layers 64 2000 >> gpurun_out/round2.log 2>&1
tail -c 7000 gpurun_out/round2.log
