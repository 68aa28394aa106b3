"""One saturated launch (9472 equal-length videos) of each DP kernel, for ncu: issue / MUFU utilisation at saturation."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import action_segmentation_b200 as pkg  # noqa: E402
from tests.helpers import random_problem, sparse_lists, to_dev  # noqa: E402

C, K, T, V = 23, 20, 1000, 9472
rng = np.random.default_rng(0)
prob = random_problem(rng, V, T, C, K, Tmin=T, chain=True, ends=True)
d = to_dev(prob)
sp = sparse_lists(prob)
g = torch.ones(V, device="cuda")
H = pkg.hsmm
for _ in range(2):
    logz, saved = H.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"], trans_pred=sp[0])
    H.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"], g, saved, trans_succ=sp[1])
    H.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"], want_score=False, trans_pred=sp[0])
torch.cuda.synchronize()
print("done")
