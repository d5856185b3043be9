#!/bin/bash
# ncu evidence for profiles/: launch list (gpu__time_duration) + full-section capture of ONE eager pass, exported to CSV
# on the box (the .ncu-rep itself is too large to bring back through gpurun_out/).
mkdir -p gpurun_out
N=${PROF_LAUNCHES:-44}
timeout 300 python tools/prof_pass.py > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/launches_${TAG:-r01}.csv python tools/prof_pass.py > gpurun_out/ncu_launches.log 2>&1 && \
timeout 1800 ncu --set full --clock-control none -s $N -c $N -o /tmp/prof_full -f python tools/prof_pass.py > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
cat gpurun_out/plain_prof.log
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${TAG:-r01}.csv 2> gpurun_out/ncu_export.log
ncu -i /tmp/prof_full.ncu-rep --page details --csv > gpurun_out/ncu_details_${TAG:-r01}.csv 2>> gpurun_out/ncu_export.log
ls -la /tmp/prof_full.ncu-rep gpurun_out/
