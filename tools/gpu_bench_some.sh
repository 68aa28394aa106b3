#!/bin/bash
# Usage under gpurun: bash tools/gpu_bench_some.sh <tag> name:args ...   (args with + instead of spaces)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for spec in "$@"; do
  name=${spec%%:*}; a=${spec#*:}; a=${a//+/ }
  timeout 900 python bench.py $a > $OUT/${TAG}_bench_${name}.json 2> $OUT/${TAG}_bench_${name}.err; echo "$name rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_${name}.json").read().strip().splitlines()[-1])
    e = d.get("e2e") or {}; c = d.get("cpu_baseline") or {}; r = d.get("roofline") or {}; s = d.get("sustained") or {}
    print("  value %.4g  ms/step %.3f  e2e %.4g  sustained %.4g  cpu %.4g  roofline %s %.3f step_frac %s binding %s" % (
        d["value"], d["ms_per_step"], e.get("value", float("nan")), s.get("value", float("nan")), c.get("value", float("nan")),
        r.get("kernel"), r.get("frac", float("nan")), r.get("step_frac"), r.get("binding")))
    print("  kernel_ms", r.get("kernel_ms"), "clocks", d.get("clocks"))
    for cell in d.get("sweep", []):
        print("   C=%d K=%d  %.4g frames/s  hbm %.3f  fp32 %.3f  %s  %s" % (cell["C"], cell["K"], cell["frames_per_s"], cell["hbm_frac"], cell["fp32_issue_frac"], cell["binding"], cell["kernel_ms"]))
except Exception as ex:
    print("  parse failed:", ex)
PY
  tail -2 $OUT/${TAG}_bench_${name}.err
done
