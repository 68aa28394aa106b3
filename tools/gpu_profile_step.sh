#!/bin/bash
# ncu evidence of the grouped bench step (configs[1]).  Usage under gpurun: bash tools/gpu_profile_step.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
KRE='regex:hsmm|etc::|wtc::|weighted_sums|dp_|emission|gen_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 190 -c 70 --csv --log-file $OUT/${TAG}_launches_step.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-sustained > $OUT/${TAG}_launches_step.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fb_kernel_grouped|vit2_kernel_grouped" -s 6 -c 2 -o /tmp/${TAG}_dp \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-sustained > $OUT/${TAG}_ncu_dp.log 2>&1
echo "ncu dp rc=$?"
ncu -i /tmp/${TAG}_dp.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_dp_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_dp.ncu-rep --page source --csv --kernel-name regex:fb_kernel > $OUT/${TAG}_ncu_dp_fb_source.csv 2>/dev/null
wc -c $OUT/${TAG}_*
