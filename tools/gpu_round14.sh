#!/bin/bash
# streaming attention bring-up: unit parity + timing per geometry (own process, short timeout), then the usual round
mkdir -p gpurun_out
LOG=gpurun_out/round14.log
echo "=== attn unit small" > $LOG
timeout 120 python tools/gpu_check.py attn 2 64 >> $LOG 2>&1; rc=$?; echo "rc=$rc" >> $LOG
if [ $rc -eq 0 ]; then
echo "=== attn unit full" >> $LOG
timeout 180 python tools/gpu_check.py attn 64 2000 >> $LOG 2>&1; rc=$?; echo "rc=$rc" >> $LOG
fi
cat $LOG
if [ $rc -ne 0 ]; then echo "attn unit failed: stopping"; exit 1; fi
TAG=${TAG:-r01s} bash tools/gpu_round10.sh
