#!/bin/bash
# one `ncu --set full` capture of ONE kernel (regex $KERNEL, launch index $SKIP) of a command, exported on the box:
# details + raw + source (SASS with stall samples) pages -> gpurun_out/<TAG>_{details,raw,source}.csv
mkdir -p gpurun_out
TAG=${TAG:-one}
CMD=${CMD:-"python tools/gpu_check.py attn 64 2000"}
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 900 ncu --set full --import-source on --clock-control none -k regex:${KERNEL:-attention_stream} -s ${SKIP:-0} -c ${COUNT:-1} -o /tmp/one -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/one.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
ncu -i /tmp/one.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/one.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
