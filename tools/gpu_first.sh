#!/bin/bash
# first GPU contact: per-part parity report (each part in its own process, each with a timeout)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for part in "stft 2 256" "stft 3 2000" "istft 2 256" "istft 3 2000" "fp32 2 256" "tcunit 2 256" "bf16 2 256" "fp32 1 64" "time_fp32 8 2000" "time_bf16 8 2000"; do
  echo "=== $part" >> gpurun_out/check.log
  timeout 300 python tools/gpu_check.py $part >> gpurun_out/check.log 2>&1
  echo "rc=$?" >> gpurun_out/check.log
done
tail -c 6000 gpurun_out/check.log
