#!/bin/bash
# A/B of an environment switch: bench.py (short) with and without $AB_VAR=1, plus the bf16 parity tests under the switch
mkdir -p gpurun_out
for v in 0 1; do
  env $AB_VAR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python -c "import json; d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('$AB_VAR=$v', d['ms_per_step'], d['e2e']['ms_per_step'], d['stage_ms'])"
done
env $AB_VAR=1 timeout 600 python -m pytest tests -q -m gpu -x -k "bf16 or full_size or enhancer" 2>&1 | tail -3
