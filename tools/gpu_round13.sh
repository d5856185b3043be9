#!/bin/bash
# cautious bring-up: full-size strip unit check first (each in its own process, short timeout), then the usual round
mkdir -p gpurun_out
LOG=gpurun_out/round13.log
echo "=== stripunit full" > $LOG
timeout 120 python tools/gpu_check.py stripunit 64 2000 >> $LOG 2>&1; rc=$?; echo "rc=$rc" >> $LOG
cat $LOG
if [ $rc -ne 0 ]; then echo "stripunit failed: stopping"; exit 1; fi
TAG=${TAG:-r01q} bash tools/gpu_round10.sh
