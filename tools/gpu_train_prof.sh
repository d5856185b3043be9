#!/bin/bash
# ncu evidence for the training step (profiles/<TAG>_train_*): launch list of one whole step + full-section capture of its GEMM /
# attention-backward kernels, summarised on the box.
TAG=${TAG:-r06}
OUT=gpurun_out/${TAG}
mkdir -p $OUT
timeout 300 python tools/time_train.py --iters 2 --mode tf32 > $OUT/time_train.json 2> $OUT/time_train.err || exit 1
cat $OUT/time_train.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/train_launches.csv python tools/time_train.py --iters 0 --no-opt --mode tf32 > $OUT/ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'cconv_tc_kernel|wgrad_tc_kernel|wgrad_small|att_bwd_du|att_bwd_w7_kernel|dec6_bwd_kernel|cconv_dgrad_cin1|lstm_train' -o /tmp/train_full -f python tools/time_train.py --iters 0 --no-opt --mode tf32 > $OUT/ncu_full.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/train_full.ncu-rep --page raw --csv > $OUT/train_raw.csv 2> $OUT/ncu_export.log
python tools/ncu_summary.py $OUT/train_raw.csv > $OUT/train_ncu_summary.csv
wc -l $OUT/train_ncu_summary.csv
rm -f $OUT/train_raw.csv
