import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import action_segmentation_b200 as pkg
from oracle import hsmm_oracle as O
from tests.helpers import random_problem, to_dev, sparse_lists
C, K = 9, 20
rng = np.random.default_rng(7 + C + K)
B, Tmax = 6, 40
prob = random_problem(rng, B, Tmax, C, K, Tmin=20, chain=True, ends=False)
prob["lengths"][1] = 3; prob["lengths"][4] = 5
end = np.full((B, C), O.BIG_NEG); end[:, C - 1] = 0.0
prob["end"] = end
prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
d = to_dev(prob)
pred, succ = sparse_lists(prob)
for name, hint in (("sparse", (pred, succ)), ("dense", (None, None))):
    for b in range(B):
        g = torch.zeros(B, device="cuda"); g[b] = 1.0
        logz, saved = pkg.hsmm.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"], trans_pred=hint[0])
        di, dt, dl, de = pkg.hsmm.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"], g, saved, trans_succ=hint[1])
        print(name, "video", b, "T", int(prob["lengths"][b]), "logz %.4g" % float(logz[b]), "d_init", np.round(di.cpu().numpy(), 4))
