#!/bin/bash
# round-2 evidence run: full -m gpu suite, smoke, default bench (with cpu baseline), reference arm, ncu launch list + full
# capture of one eager pass of the complex (dcs) plan and of the real (drs) plan, exported as summaries for profiles/.
mkdir -p gpurun_out
TAG=${TAG:-r04a}
LOG=gpurun_out/${TAG}_evidence.log
echo "=== pytest gpu" > $LOG
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 2>&1 | tail -6 >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
echo "=== bench default" >> $LOG
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "rc=$?" >> $LOG
echo "=== bench reference arm" >> $LOG
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "rc=$?" >> $LOG
cat $LOG
N=${PROF_LAUNCHES:-43}
timeout 300 python tools/prof_pass.py > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/${TAG}_launches.csv python tools/prof_pass.py > gpurun_out/ncu_launches.log 2>&1 && \
timeout 1800 ncu --set full --clock-control none -s $N -c $N -o /tmp/prof_full -f python tools/prof_pass.py > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/plain_prof.log
ncu -i /tmp/prof_full.ncu-rep --page raw --csv > /tmp/ncu_raw_${TAG}.csv 2> gpurun_out/ncu_export.log
python tools/ncu_summary.py /tmp/ncu_raw_${TAG}.csv > gpurun_out/${TAG}_ncu_summary.csv
if [ -n "$REAL_LAUNCHES" ]; then
  N=$REAL_LAUNCHES
  PROF_VARIANT=drs timeout 300 python tools/prof_pass.py > gpurun_out/plain_prof_real.log 2>&1 && \
  PROF_VARIANT=drs timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/${TAG}_real_launches.csv python tools/prof_pass.py > gpurun_out/ncu_launches_real.log 2>&1 && \
  PROF_VARIANT=drs timeout 1800 ncu --set full --clock-control none -s $N -c $N -o /tmp/prof_full_real -f python tools/prof_pass.py > gpurun_out/ncu_full_real.log 2>&1
  echo "ncu real rc=$?"; cat gpurun_out/plain_prof_real.log
  ncu -i /tmp/prof_full_real.ncu-rep --page raw --csv > /tmp/ncu_raw_${TAG}_real.csv 2>> gpurun_out/ncu_export.log
  python tools/ncu_summary.py /tmp/ncu_raw_${TAG}_real.csv > gpurun_out/${TAG}_real_ncu_summary.csv
fi
ls -la gpurun_out/${TAG}_*
