import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import action_segmentation_b200 as pkg
from oracle import hsmm_oracle as O
from tests.helpers import random_problem, to_dev, sparse_lists
f32 = lambda x: None if x is None else x.astype(np.float32).astype(np.float64)
rng = np.random.default_rng(33)
B, Tmax, C, K = 2, 3000, 23, 20
prob = random_problem(rng, B, Tmax, C, K, Tmin=2500, chain=True, ends=True, scale=2.5)
prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
d = to_dev(prob); sp = sparse_lists(prob)
w = np.ones(B)
ref_logz, acc = O.batch_logz_and_counts(f32(prob["em"]), prob["lengths"], f32(prob["init"]), f32(prob["trans"]), f32(prob["lenp"]), prob["end"], w)
for name, lin, gen in [("lin", True, False), ("logdomain", False, False), ("general", False, True)]:
    pkg._lib.set_linear_window(lin); pkg._lib.set_generic_dp(gen)
    logz, saved = pkg.hsmm.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"], trans_pred=sp[0])
    g = torch.ones(B, device="cuda")
    di, dt, dl, de = pkg.hsmm.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"], g, saved, trans_succ=sp[1])
    de = de[:, :, :C].cpu().numpy().astype(np.float64)
    print(name, "logz err", np.abs(logz.cpu().numpy() - ref_logz).max())
    for b in range(B):
        T = int(prob["lengths"][b])
        rowsum = de[b, :T].sum(axis=1) - 1.0
        err = np.abs(de[b, :T] - acc["E_em"][b, :T]).max(axis=1)
        idx = [0, 1, 10, 100, 500, 1000, 1500, 2000, T - 100, T - 10, T - 1]
        print("  b", b, "rowsum-1 at", idx, ["%.1e" % rowsum[i] for i in idx])
        print("       max err per frame", ["%.1e" % err[i] for i in idx], "overall", "%.2e" % err.max(), "argmax", int(err.argmax()))
    print("  E_trans err", np.abs(dt.cpu().numpy() - acc["E_trans"]).max() / np.abs(acc["E_trans"]).max(), "E_len", np.abs(dl.cpu().numpy() - acc["E_len"]).max() / np.abs(acc["E_len"]).max(),
          "E_init", np.abs(di.cpu().numpy() - acc["E_init"]).max())
