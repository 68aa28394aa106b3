"""Run one task's emission / weighted-sums launches a few times (for ncu --set full --import-source on)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from action_segmentation_b200 import hsmm  # noqa: E402

sys.argv = [sys.argv[0]]
args = bench.parse()
gen = torch.Generator().manual_seed(1)
tk = bench.make_task(0, 11, 128, 200, 20, 1000, 3000, False, gen, "cuda:0")
for _ in range(3):
    em, rowterm, offset = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, None, tk.lengths_i32, params=tk.eparams)
    w = torch.softmax(em, dim=-1).contiguous()
    hsmm.weighted_feature_sums(tk.X, w, tk.C, tk.lengths_i32)
torch.cuda.synchronize()
print("done")
