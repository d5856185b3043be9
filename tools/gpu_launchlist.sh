#!/bin/bash
# quick per-kernel device times of one eager pass (cold-cache, serialised): ncu launch list only
mkdir -p gpurun_out
N=${PROF_LAUNCHES:-44}
timeout 300 python tools/prof_pass.py > gpurun_out/plain_prof.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/launches_${TAG:-x}.csv python tools/prof_pass.py > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"; cat gpurun_out/plain_prof.log
python - <<'PY'
import csv,os,collections
tag=os.environ.get("TAG","x")
rows=list(csv.reader(open(f"gpurun_out/launches_{tag}.csv")))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    k=r[4].replace("dcs::","").split("(")[0][:50]; agg.setdefault(k,[]).append(float(r[-1])/1e3)
for k,v in agg.items(): print(f"{k:52s} n={len(v):2d} total_us={sum(v):9.1f}  each={' '.join(f'{x:.0f}' for x in v[:14])}")
PY
