#!/bin/bash
# GPU-box pass: parity tests (+ optional bench line).  Usage under gpurun: bash tools/gpu_tests.sh <tag> [pytest args]
TAG=${1:-r02}
shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
timeout 2400 python -m pytest tests -m gpu -q "$@" > $OUT/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
tail -25 $OUT/${TAG}_pytest_gpu.log
