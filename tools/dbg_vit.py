"""fp64 score gap between the decoded path and the oracle's best path, per video (near-tie diagnosis)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import action_segmentation_b200 as pkg  # noqa: E402
from oracle import hsmm_oracle as O  # noqa: E402
from tests.helpers import random_problem, sparse_lists, to_dev  # noqa: E402

B, Tmax, C, K, scale = 16, 400, 23, 20, 3.0
rng = np.random.default_rng(500 + C * 7 + K)
prob = random_problem(rng, B, Tmax, C, K, Tmin=1, chain=True, ends=True, scale=scale)
prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
d = to_dev(prob)
sp = sparse_lists(prob)
for mode in (True, False):
    pkg._lib.set_linear_window(mode)
    s1, l1, sc1 = pkg.hsmm.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"],
                                          trans_pred=sp[0])
    spans = s1.cpu().numpy()
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    for b in range(B):
        T = int(prob["lengths"][b])
        if T < C:
            continue
        best, segs = O.viterbi(f32(prob["em"][b, :T]), f32(prob["init"]), f32(prob["trans"]), f32(prob["lenp"]), prob["end"][b])
        mine = O.segments_from_spans(spans[b], T)
        s = O.path_score(mine, f32(prob["em"][b, :T]), f32(prob["init"]), f32(prob["trans"]), f32(prob["lenp"]), prob["end"][b])
        same = (O.segs_to_spans(segs, T, C, Tmax + 1) == spans[b]).all()
        print("lin" if mode else "old", b, T, "same" if same else "DIFF", "gap %.3e" % (best - s), "best %.2f" % best, "nseg", len(segs), len(mine))
pkg._lib.set_linear_window(True)
