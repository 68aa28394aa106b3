#!/bin/bash
# One GPU-box pass: parity tests, bench line, reference arm, ncu launch list, ncu full capture of the step's kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests] [sat]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
  tail -3 $OUT/${TAG}_pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hsmm|etc::|wtc::|weighted_sums|dp_|emission' -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $OUT/${TAG}_launches_bench.log 2>&1
echo "ncu launches rc=$?"
# full capture of one task's kernels inside the bench step; summarised to CSV on the box (the .ncu-rep is too large to bring back)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dp_lin|dp_vit2|emission_tc|weighted_sums' -s 10 -c 10 -o /tmp/${TAG}_bench_top \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu bench rc=$?"
ncu -i /tmp/${TAG}_bench_top.ncu-rep --page raw --csv > $OUT/${TAG}_bench_top_raw.csv 2>/dev/null
for k in emission_tc weighted_sums_tc; do
  ncu -i /tmp/${TAG}_bench_top.ncu-rep --page source --csv --kernel-name regex:$k > $OUT/${TAG}_src_$k.csv 2>/dev/null
done
if [ "$3" == "sat" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dp_lin|dp_vit2' -s 3 -c 3 -o /tmp/${TAG}_dp_sat \
    python tools/sat_profile.py > $OUT/${TAG}_ncu_sat.log 2>&1
  ncu -i /tmp/${TAG}_dp_sat.ncu-rep --page raw --csv > $OUT/${TAG}_dp_sat_raw.csv 2>/dev/null
fi
ls -la $OUT
