#!/bin/bash
# bench lines of the non-default BASELINE configurations -> gpurun_out/lines_<TAG>/ (copied to profiles/bench_lines/)
TAG=${TAG:-r2}
OUT=gpurun_out/lines_${TAG}
mkdir -p $OUT
N=${NGPU:-1}
run() {  # name, args...
  name=$1; shift
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > $OUT/${name}_n1.json 2> $OUT/${name}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" > $OUT/${name}_n${N}.json 2> $OUT/${name}_n${N}.err
  fi
  echo "$name rc=$?"; tail -c 600 $OUT/${name}_n${N}.json | head -c 600; echo; tail -3 $OUT/${name}_n${N}.err
}
for what in ${WHAT:-dcs dc longform}; do
  case $what in
    dcs) run dcs --steps 20 --warmup 3 --no-cpu-baseline ;;
    dc) run dc --variant dc --steps 20 --warmup 3 --no-cpu-baseline ;;
    dr) run dr --variant dr --steps 20 --warmup 3 --no-cpu-baseline ;;
    drs) run drs --variant drs --steps 20 --warmup 3 --no-cpu-baseline ;;
    longform) run longform --workload longform --steps 5 --warmup 2 --no-cpu-baseline ;;
  esac
done
