#!/bin/bash
# ncu evidence for the round: launch list of the bench step, --set full of the step's kernels, per-config top kernels,
# compute-sanitizer racecheck/memcheck on the DP kernels.  Usage under gpurun: bash tools/gpu_profile.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out; mkdir -p $OUT
KRE='regex:hsmm|etc::|wtc::|weighted_sums|dp_|emission|gen_|moments|gold|onehot'
# 1. every launch of one bench step (configs[1]) with its device time
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 800 --csv --log-file $OUT/${TAG}_launches_cfg1.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-sustained > $OUT/${TAG}_launches_cfg1.log 2>&1
echo "launch list rc=$?"
# 1b. --set full of the grouped DP kernels and the streaming kernels inside the bench step
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:_grouped|emission_tc|weighted_sums_tc" -s 40 -c 12 -o /tmp/${TAG}_step \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph --no-sustained > $OUT/${TAG}_ncu_step.log 2>&1
echo "ncu step rc=$?"
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_step_raw.csv 2>/dev/null
# 2. --set full of one task's kernels per config (summarised to CSV on the box; the .ncu-rep is too large to bring back)
for spec in "4:133:200:512:2:3" "4:64:100:1024:2:3" "4:16:50:2048:2:3"; do
  IFS=: read cfg C K V SKIP CNT <<< "$spec"
  PROFILE_CONFIG=$cfg PROFILE_C=$C PROFILE_K=$K PROFILE_VIDEOS=$V timeout 900 ncu --set full --clock-control none --import-source on \
    -k "$KRE" -s $SKIP -c $CNT -o /tmp/${TAG}_cfg${cfg}_C${C}_K${K} python profiles/profile_one_task.py > $OUT/${TAG}_ncu_cfg${cfg}_C${C}_K${K}.log 2>&1
  echo "ncu cfg$cfg C=$C K=$K rc=$?"
  ncu -i /tmp/${TAG}_cfg${cfg}_C${C}_K${K}.ncu-rep --page raw --csv > $OUT/${TAG}_ncu_cfg${cfg}_C${C}_K${K}_raw.csv 2>/dev/null
done
# 3. sanitizer: racecheck + memcheck on the DP fast paths and the general kernels (small shapes)
for tool in racecheck memcheck; do
  timeout 1200 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py tests/test_gpu_general_fullsize.py -m gpu -q -x \
    -k "linear_window_matches or deferred_argmax or (general_kernels_vs_oracle and C23) or (logz_and_counts_random and C23_K20)" \
    > $OUT/${TAG}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?"; tail -4 $OUT/${TAG}_sanitizer_${tool}.log
done
ls -la $OUT | tail -30
