"""Markdown table of the committed bench lines (profiles/<tag>_bench_*.json) for DESIGN.md section 6.
Usage: python tools/make_results_table.py r02k"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = [("cfg0", "configs[0] S6-shape C=11 K=100 dense"), ("cfg1", "configs[1] U7-shape, 18 tasks, K=20 chain (driver default)"),
        ("cfg2", "configs[2] = [1] + narration penalty"), ("cfg3_K200", "configs[3] Breakfast C=48 D=64 K=200"),
        ("cfg3_K500", "configs[3] K=500 (general kernels)"), ("cfg4", "configs[4] decode sweep, 9 cells")]
print("| config | frames/s (device) | ms/step | sustained ≥2 s | e2e frames/s (host buffers) | reference on host cores | dominant call: HBM frac (alg. bytes) | step HBM frac | FP32-issue frac | binding | file |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for key, name in rows:
    path = os.path.join(ROOT, "profiles", "%s_bench_%s.json" % (tag, key))
    if not os.path.exists(path):
        continue
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r, e, s, c = d.get("roofline") or {}, d.get("e2e") or {}, d.get("sustained") or {}, d.get("cpu_baseline") or {}
    comp = (r.get("compute_ceiling") or {}).get("frac")
    f = lambda v, fmt="%.3g": "—" if v is None else fmt % v  # noqa: E731
    print("| %s | %s | %s | %s | %s | %s (%s cores) | %s %s | %s | %s | %s | `profiles/%s` |" % (
        name, f(d["value"], "%.4g"), f(d["ms_per_step"], "%.3f"), f(s.get("value"), "%.4g"), f(e.get("value"), "%.4g"),
        f(c.get("value"), "%.4g"), c.get("cores", "—"), r.get("kernel", "—"), f(r.get("frac"), "%.3f"), f(r.get("step_frac"), "%.3f"),
        f(comp, "%.3f"), r.get("binding", "—"), os.path.basename(path)))
    if d.get("sweep"):
        print()
        print("| C | K | frames/s | ms | emission ms | Viterbi ms | HBM frac | FP32-issue frac | binding | Viterbi kernel |")
        print("|---|---|---|---|---|---|---|---|---|---|")
        for cell in d["sweep"]:
            print("| %d | %d | %.4g | %.1f | %.1f | %.1f | %.3f | %.3f | %s | %s |" % (
                cell["C"], cell["K"], cell["frames_per_s"], cell["ms_per_step"], cell["kernel_ms"]["emission"], cell["kernel_ms"]["viterbi"],
                cell["hbm_frac"], cell["fp32_issue_frac"], cell["binding"], cell["viterbi_variant"].split("/")[0]))
