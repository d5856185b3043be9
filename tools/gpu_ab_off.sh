#!/bin/bash
# A/B of an environment switch that DISABLES a new path: bench.py (short) with $AB_VAR=0 (old) and unset (new), after the tests in $AB_TESTS
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -k "${AB_TESTS:-attention}" 2>&1 | tail -5
for v in 0 1; do
  env $AB_VAR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${AB_ARGS} > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('$AB_VAR=$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
done
