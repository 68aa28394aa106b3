#!/bin/bash
# ncu evidence of the grouped bench step (configs[1]).  Usage under gpurun: bash tools/gpu_profile_step.sh <tag>
# 1. launch list of two steps (gpu__time_duration per launch: cold-cache and serialised -> compare SHARES with bench.py);
# 2. --set full of the step's four kernel families (one launch each), exported as raw CSV (+ source page of the DP kernel);
# 3. the step timeline from CUPTI kernel records (tools/step_timeline.py).
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
KRE='regex:hsmm|etc::|wtc::|weighted_sums|dp_|emission|gen_|upload'
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-sustained"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 120 -c 200 --csv --log-file $OUT/${TAG}_launches_step.csv \
  $B > $OUT/${TAG}_launches_step.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:fb_kernel_grouped|vit2_kernel_grouped" -s 6 -c 2 -o /tmp/${TAG}_dp \
  $B > $OUT/${TAG}_ncu_dp.log 2>&1
echo "ncu dp rc=$?"
ncu -i /tmp/${TAG}_dp.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_dp_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_dp.ncu-rep --page source --csv --kernel-name regex:fb_kernel > $OUT/${TAG}_ncu_dp_fb_source.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k "regex:emission_tc_kernel" -s 58 -c 2 -o /tmp/${TAG}_em \
  $B > $OUT/${TAG}_ncu_em.log 2>&1
echo "ncu emission rc=$?"
ncu -i /tmp/${TAG}_em.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_em_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k "regex:weighted_sums_tc_kernel" -s 58 -c 2 -o /tmp/${TAG}_ws \
  $B > $OUT/${TAG}_ncu_ws.log 2>&1
echo "ncu wsums rc=$?"
ncu -i /tmp/${TAG}_ws.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_ws_raw.csv 2>/dev/null
python tools/step_timeline.py --out $OUT/${TAG}_step_timeline.txt > /dev/null 2>&1
echo "timeline rc=$?"
wc -c $OUT/${TAG}_*
