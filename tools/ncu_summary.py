"""Summarise an .ncu-rep (read here, no GPU): one row per profiled launch with the metrics the roofline argument uses.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_x_summary.csv [--traffic profiles/ncu_traffic.json]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio"]
NAMES = {"dp_backward": "logz_backward", "dp_forward_kernel<(bool)0": "logz_forward", "dp_forward_kernel<(bool)1": "viterbi",
         "dp_lin_forward": "logz_forward", "dp_lin_backward": "logz_backward", "dp_vit2": "viterbi",
         "emission_tc": "emission", "emission_kernel": "emission", "weighted_sums": "weighted_feature_sums"}


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * mult


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):  # already exported on the GPU box (ncu -i x.ncu-rep --page raw --csv)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in ix]
    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keys)
        w.writerow([units[ix[k]] for k in keys])
        for r in rows[2:]:
            w.writerow([r[ix[k]][:120] for k in keys])
            name = r[ix["Kernel Name"]]
            for pat, nm in NAMES.items():
                if pat in name:
                    t = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + \
                        to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
                    traffic.setdefault(nm, []).append(t)
    if "--traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic") + 1]
        # one API call = the main kernel + (for the DP) the near-empty only-flagged fallback launch behind it:
        # bytes per call = total bytes / number of launches that moved more than 1 MB
        json.dump({k: sum(v) / max(1, sum(1 for t in v if t > 1e6)) for k, v in traffic.items()} |
                  {"_note": "mean dram bytes (read+write) per profiled API call, from " + rep}, open(path, "w"), indent=1)
    print("wrote", out, {k: len(v) for k, v in traffic.items()})


if __name__ == "__main__":
    main()
