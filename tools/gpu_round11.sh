#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/round11.log
echo "=== stripunit small" > $LOG
timeout 120 python tools/gpu_check.py stripunit 2 256 >> $LOG 2>&1; echo "rc=$?" >> $LOG
echo "=== stripunit full" >> $LOG
timeout 300 python tools/gpu_check.py stripunit 64 2000 >> $LOG 2>&1; echo "rc=$?" >> $LOG
cat $LOG
TAG=${TAG:-r01h} bash tools/gpu_round10.sh
