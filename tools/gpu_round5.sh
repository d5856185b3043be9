#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest gpu" > gpurun_out/round5.log
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -30 >> gpurun_out/round5.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/round5.log
echo "=== layers" >> gpurun_out/round5.log
timeout 300 python tools/gpu_check.py layers 64 2000 2>&1 | grep -v '"layer"' >> gpurun_out/round5.log
echo "=== bench ours" >> gpurun_out/round5.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err; echo "rc=$?" >> gpurun_out/round5.log
cat gpurun_out/bench5.json >> gpurun_out/round5.log; tail -5 gpurun_out/bench5.err >> gpurun_out/round5.log
echo "=== ncu" >> gpurun_out/round5.log
timeout 300 python tools/prof_pass.py > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cconv_tc_kernel -s 21 -c 21 -o gpurun_out/prof_tc python tools/prof_pass.py > gpurun_out/ncu_tc.log 2>&1
echo "ncu tc rc=$?" >> gpurun_out/round5.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/round5.log
ls -la gpurun_out >> gpurun_out/round5.log
tail -c 6000 gpurun_out/round5.log
