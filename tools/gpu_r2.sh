#!/bin/bash
# round-2 GPU check: -m gpu suite, smoke(), bench in the tensor-core modes.  TAG names the outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-r2a}
LOG=gpurun_out/${TAG}.log
echo "=== pytest gpu" > $LOG
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 ${PYTEST_ARGS:--x} 2>&1 | tail -${PYTEST_TAIL:-40} >> $LOG
echo "=== smoke" >> $LOG
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $LOG 2>&1
for MODE in ${MODES:-fp16 bf16}; do
  echo "=== bench $MODE" >> $LOG
  timeout 600 python bench.py --mode $MODE --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_${MODE}.json 2> gpurun_out/bench_${TAG}_${MODE}.err; echo "rc=$?" >> $LOG
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${TAG}_${MODE}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['stage_ms'])" >> $LOG 2>&1
  tail -5 gpurun_out/bench_${TAG}_${MODE}.err >> $LOG
done
cat $LOG
