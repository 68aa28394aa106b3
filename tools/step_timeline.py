"""Kernel timeline of ONE configs[1] step (the CUDA graph bench.py replays), from torch.profiler's CUPTI kernel records:
per kernel family the first start, the last end and the busy time, relative to the step's first kernel.  It answers
"what overlaps what" for the step's streams; the absolute durations carry the profiler's overhead.

    python tools/step_timeline.py [--config 1] [--out gpurun_out/timeline.txt]
Environment switches of bench.py (HSMM_BENCH_GROUPS, HSMM_BENCH_FUSED, HSMM_BENCH_VIT_LAST, HSMM_BENCH_PRIO) apply.
"""
import argparse
import os
import re
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def family(name):
    for key, fam in (("emission", "emission"), ("weighted_sums", "wsums"), ("vit", "viterbi"), ("fb_kernel", "fwd+bwd"),
                     ("forward", "forward"), ("backward", "backward"), ("nccl", "nccl")):
        if key in name:
            return fam
    return re.sub(r"<.*", "", name.split("(")[0])[-40:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    sys.argv = [sys.argv[0], "--config", str(a.config)]
    args = bench.parse()
    cfg = bench.config_of(args)
    device = torch.device("cuda:0")
    torch.cuda.set_device(device)
    tasks = bench.make_workload(args, cfg, 0, device)
    layout, total = bench.packed_layout(tasks)
    packed = torch.zeros(total, device=device)
    fused = os.environ.get("HSMM_BENCH_FUSED", "0" if cfg["narration"] else "1") == "1"
    n_groups = int(os.environ.get("HSMM_BENCH_GROUPS", "1" if fused else "3"))
    streams = bench.make_streams(len(tasks))

    def step():
        return bench.device_step_grouped(tasks, streams, packed, layout, 1, reduce=False, n_groups=n_groups, fused=fused)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outs = step()  # noqa: F841
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            graph.replay()
            torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    # the host synchronises between replays: the steps do not interleave and launch the same number of kernels
    per = len(ev) // a.steps
    steps = [ev[i * per:(i + 1) * per] for i in range(a.steps)]
    lines = []
    for si, st in enumerate(steps[-2:]):
        t0 = min(e.time_range.start for e in st)
        t1 = max(e.time_range.end for e in st)
        lines.append(f"step {si}: {len(st)} kernels, {t1 - t0:.0f} us")
        fams = {}
        for e in st:
            f = fams.setdefault(family(e.name), [1e30, 0, 0.0, 0])
            f[0] = min(f[0], e.time_range.start - t0)
            f[1] = max(f[1], e.time_range.end - t0)
            f[2] += e.time_range.end - e.time_range.start
            f[3] += 1
        for k, f in sorted(fams.items(), key=lambda kv: kv[1][0]):
            lines.append(f"  {k:28s} n={f[3]:3d}  first start {f[0]:8.0f}  last end {f[1]:8.0f}  sum of durations {f[2]:8.0f} us")
    text = "\n".join(lines)
    print(text)
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(text + "\n")


if __name__ == "__main__":
    main()
