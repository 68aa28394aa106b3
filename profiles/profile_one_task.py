"""One CrossTask-shaped task (C=23, K=20, D=200) through every kernel of the hot path, for ncu.
    ncu --set full --clock-control none --import-source on -k regex:dp_ -c 4 -o gpurun_out/prof python profiles/profile_one_task.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from action_segmentation_b200 import hsmm  # noqa: E402

V = int(os.environ.get("PROFILE_VIDEOS", "2048"))
steps = int(os.environ.get("PROFILE_STEPS", "11"))
K = int(os.environ.get("PROFILE_K", "20"))
gen = torch.Generator().manual_seed(1)
tk = bench.make_task(0, steps, V, 200, K, 1000, 3000, False, gen, "cuda")
for it in range(2):
    em, rowterm, offset = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32)
    sparse = os.environ.get('PROFILE_DENSE', '0') != '1'
    pred, succ = (tk.pred, tk.succ) if sparse else (None, None)
    logz, saved = hsmm.logz_forward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order, trans_pred=pred)
    g = torch.full((tk.V,), 1.0 / tk.V, device="cuda")
    d = hsmm.logz_backward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, tk.lengths_i32, tk.order, g, saved, trans_succ=succ)
    hsmm.weighted_feature_sums(tk.X, d[3], tk.C, tk.lengths_i32)
    hsmm.viterbi_decode(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order, tk.class_ids, trans_pred=pred)
    torch.cuda.synchronize()
print("frames", tk.frames, "logz mean", float(logz.mean()))
