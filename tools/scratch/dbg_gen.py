import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import action_segmentation_b200 as pkg
from oracle import hsmm_oracle as O
from tests.helpers import random_problem, to_dev
pkg._lib.set_generic_dp(True)
for (B, Tmax, C, K, chain) in [(1, 12, 4, 40, False), (1, 12, 4, 40, True), (2, 30, 3, 6, False)]:
    rng = np.random.default_rng(1)
    prob = random_problem(rng, B, Tmax, C, K, Tmin=Tmax, chain=chain, ends=chain)
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    logz, saved = pkg.hsmm.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"])
    g = torch.ones(B, device="cuda")
    di, dt, dl, de = pkg.hsmm.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"], g, saved)
    ref_logz, acc = O.batch_logz_and_counts(prob["em"], prob["lengths"], prob["init"], prob["trans"], prob["lenp"], prob["end"])
    print("case", B, Tmax, C, K, chain)
    print(" logz", logz.cpu().numpy(), ref_logz)
    print(" d_init", di.cpu().numpy(), acc["E_init"])
    print(" d_em[0,:3]", de[0, :3, :C].cpu().numpy(), acc["E_em"][0, :3])
    print(" d_len sum", float(dl.sum()), acc["E_len"].sum(), " d_trans sum", float(dt.sum()), acc["E_trans"].sum())
