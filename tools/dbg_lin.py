"""Which videos do the linear-window kernels hand to the log-domain kernels, and why (bflag reason bits)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import action_segmentation_b200 as pkg  # noqa: E402
from tests.helpers import random_problem, sparse_lists, to_dev  # noqa: E402
from tests.test_gpu_parity import _saved_flags  # noqa: E402

C, K, T, V = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (23, 20, 2000, 128)
rng = np.random.default_rng(0)
prob = random_problem(rng, V, T, C, K, Tmin=T, chain=True, ends=True)
d = to_dev(prob)
sp = sparse_lists(prob)
H = pkg.hsmm
logz, saved = H.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"], trans_pred=sp[0])
ff = _saved_flags(saved, V, T, C)[0].copy()
g = torch.ones(V, device="cuda")
H.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"], g, saved, trans_succ=sp[1])
torch.cuda.synchronize()
bf = _saved_flags(saved, V, T, C)[1]
print("fflag histogram", np.unique(ff, return_counts=True))
print("bflag histogram", np.unique(bf, return_counts=True))
print("rates", np.exp(prob["lenp"][1] - prob["lenp"][0] if False else 0))
