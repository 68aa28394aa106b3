"""Time the two feature-streaming kernels (emission scoring, class-weighted feature sums) on one bench task:
CUDA events around back-to-back launches over several tasks' worth of features (> L2), microseconds and GB/s of X.
Usage: python tools/time_stream_kernels.py [steps_in_task]   (experiment switches come from the environment)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from action_segmentation_b200 import hsmm  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 11
videos = int(sys.argv[2]) if len(sys.argv) > 2 else 128
sys.argv = [sys.argv[0]]
args = bench.parse()
gen = torch.Generator().manual_seed(1)
tasks = [bench.make_task(i, steps, videos, 200, 20, 1000, 3000, False, gen, "cuda:0") for i in range(4 if videos <= 128 else 2)]  # > L2 in total
ws = []
for tk in tasks:
    em, rowterm, offset = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, None, tk.lengths_i32, params=tk.eparams)
    ws.append(torch.softmax(em, dim=-1).contiguous())
torch.cuda.synchronize()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i, tk in enumerate(tasks):
            fn(i, tk)
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b) / len(tasks))
    return best


frames = sum(tk.frames for tk in tasks) / len(tasks)
xbytes = frames * 200 * 4
t_e = timed(lambda i, tk: hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, None, tk.lengths_i32, params=tk.eparams))
t_w = timed(lambda i, tk: hsmm.weighted_feature_sums(tk.X, ws[i], tk.C, tk.lengths_i32))
print("C=%d frames/task=%d  emission %.1f us (%.0f GB/s of X)  weighted sums %.1f us (%.0f GB/s of X)  env=%s" % (
    tasks[0].C, frames, t_e * 1e3, xbytes / t_e / 1e6, t_w * 1e3, xbytes / t_w / 1e6,
    {k: v for k, v in os.environ.items() if k.startswith("HSMM_")}))
