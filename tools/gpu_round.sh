#!/bin/bash
# GPU round: smoke, pytest -m gpu, bench (both arms), ncu launch list.  Logs under gpurun_out/.
mkdir -p gpurun_out
echo "=== smoke" > gpurun_out/round.log
timeout 300 python __graft_entry__.py smoke >> gpurun_out/round.log 2>&1; echo "rc=$?" >> gpurun_out/round.log
echo "=== pytest gpu" >> gpurun_out/round.log
timeout 900 python -m pytest tests -q -m gpu -x --timeout 600 2>&1 | tail -40 >> gpurun_out/round.log; echo "rc=${PIPESTATUS[0]}" >> gpurun_out/round.log
echo "=== bench ours" >> gpurun_out/round.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "rc=$?" >> gpurun_out/round.log
cat gpurun_out/bench_ours.json >> gpurun_out/round.log; tail -5 gpurun_out/bench_ours.err >> gpurun_out/round.log
echo "=== bench fp32" >> gpurun_out/round.log
timeout 600 python bench.py --steps 5 --warmup 3 --mode fp32 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "rc=$?" >> gpurun_out/round.log
cat gpurun_out/bench_fp32.json >> gpurun_out/round.log; tail -5 gpurun_out/bench_fp32.err >> gpurun_out/round.log
echo "=== bench reference" >> gpurun_out/round.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?" >> gpurun_out/round.log
cat gpurun_out/bench_ref.json >> gpurun_out/round.log
tail -c 5000 gpurun_out/round.log
