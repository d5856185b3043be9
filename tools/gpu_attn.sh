#!/bin/bash
# streaming attention: unit parity + timing per geometry (small, then full size), then one ncu capture of the C = 8 launch
mkdir -p gpurun_out
LOG=gpurun_out/attn.log
echo "=== attn unit small" > $LOG
timeout 120 python tools/gpu_check.py attn 2 64 >> $LOG 2>&1; rc=$?; echo "rc=$rc" >> $LOG
if [ $rc -eq 0 ]; then
echo "=== attn unit full" >> $LOG
timeout 180 python tools/gpu_check.py attn 64 2000 >> $LOG 2>&1; rc=$?; echo "rc=$rc" >> $LOG
fi
cat $LOG
if [ $rc -eq 0 ] && [ -n "$NCU" ]; then TAG=${TAG:-as2} KERNEL=attention_stream SKIP=1 bash tools/gpu_ncu_one.sh; fi
