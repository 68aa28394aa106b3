#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu full capture of the DP kernels.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
cat $OUT/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hsmm|etc::|weighted_sums|dp_|emission' -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $OUT/${TAG}_launches_bench.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dp_' -s 3 -c 3 -o $OUT/${TAG}_dp_sat \
  python tools/sat_profile.py > $OUT/${TAG}_ncu_sat.log 2>&1
echo "ncu sat rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dp_|emission_tc|weighted' -s 15 -c 10 -o $OUT/${TAG}_bench_top \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu bench rc=$?"
ls -la $OUT
