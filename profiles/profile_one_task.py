"""One task of a BASELINE config through every kernel of the hot path, for ncu.
    PROFILE_CONFIG=1 PROFILE_VIDEOS=2048 ncu --set full --clock-control none --import-source on -k regex:dp_ -c 4 \
        -o gpurun_out/prof python profiles/profile_one_task.py
PROFILE_CONFIG: 0 (C=11,K=100 dense), 1 (C=23,K=20 chain), 2 (= 1 + narration), 3 (Breakfast C=48,D=64,K=PROFILE_K or 200),
4 (decode sweep cell C=PROFILE_C/K=PROFILE_K)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from action_segmentation_b200 import hsmm  # noqa: E402

cfgno = int(os.environ.get("PROFILE_CONFIG", "1"))
V = int(os.environ.get("PROFILE_VIDEOS", "2048"))
sys.argv = [sys.argv[0], "--config", str(cfgno)]
args = bench.parse()
cfg = bench.config_of(args)
if os.environ.get("PROFILE_K"):
    cfg["K"] = int(os.environ["PROFILE_K"])
C = int(os.environ.get("PROFILE_C", {0: 11, 1: 23, 2: 23, 3: 48, 4: 133}[cfgno]))
gen = torch.Generator().manual_seed(1)
tk = bench.make_task(C, cfg, gen, "cuda", V=V)
decode_only = cfgno == 4
for it in range(2):
    em, rowterm, offset = hsmm.emission_scores(tk.X, tk.means, tk.cov_diag, tk.penalty, tk.lengths_i32)
    if not decode_only:
        xp = tk.penalty is not None
        logz, saved = hsmm.logz_forward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order,
                                        trans_pred=tk.pred, f64_state=xp)
        d = hsmm.logz_backward(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, tk.lengths_i32, tk.order, tk.gradw, saved,
                               trans_succ=tk.succ, f64_state=xp)
        hsmm.weighted_feature_sums(tk.X, d[3], tk.C, tk.lengths_i32)
    hsmm.viterbi_decode(em, tk.C, tk.init, tk.trans, tk.lenp, tk.end, offset, tk.lengths_i32, tk.order, tk.class_ids, trans_pred=tk.pred)
    torch.cuda.synchronize()
print("frames", tk.frames, "config", cfgno, "C", C, "K", cfg["K"], "videos", V)
