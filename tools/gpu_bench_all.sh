#!/bin/bash
# All BASELINE configs on one box.  Usage under gpurun: bash tools/gpu_bench_all.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
run() { name=$1; shift; timeout 900 python bench.py "$@" > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err; echo "$name rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}; c = d.get("cpu_baseline") or {}; r = d.get("roofline") or {}; s = d.get("sustained") or {}
    print("  value %.4g  ms/step %.3f  e2e %.4g  sustained %.4g  cpu %.4g  roofline %s %.3f step_frac %s binding %s" % (
        d["value"], d["ms_per_step"], e.get("value", float("nan")), s.get("value", float("nan")), c.get("value", float("nan")),
        r.get("kernel"), r.get("frac", float("nan")), r.get("step_frac"), r.get("binding")))
    for cell in d.get("sweep", []):
        print("   C=%d K=%d  %.4g frames/s  hbm %.3f  fp32 %.3f  %s  %s" % (cell["C"], cell["K"], cell["frames_per_s"], cell["hbm_frac"], cell["fp32_issue_frac"], cell["binding"], cell["kernel_ms"]))
except Exception as ex:
    print("  parse failed:", ex)
PY
tail -3 $OUT/${TAG}_bench_${name}.err; }
run cfg1
run cfg0 --config 0
run cfg2 --config 2
run cfg3_K200 --config 3
run cfg3_K500 --config 3 --max-span 500 --steps 2 --warmup 1
run cfg4 --config 4
run ref_cfg1 --impl reference --steps 2 --warmup 1
