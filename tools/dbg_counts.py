"""Debug: expected-count errors vs the fp64 oracle for chain + narration problems at several penalty weights."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hsmm_oracle as O
from tests.helpers import random_problem, rel_err, sparse_lists, to_dev
import action_segmentation_b200 as pkg

def run(shape, pen, sparse=True, seed=None):
    B, Tmax, C, K, chain, ends = shape
    rng = np.random.default_rng(200 + C * 7 + K if seed is None else seed)
    prob = random_problem(rng, B, Tmax, C, K, Tmin=1, chain=chain, ends=ends, narration=False)
    if pen:
        em = prob["em"]
        for b in range(B):
            for c in range(1, C, 2):
                lo = int(rng.integers(0, max(1, prob["lengths"][b]))); hi = lo + int(rng.integers(2, 12))
                p = np.full(Tmax, -pen); p[lo:hi] = 0.0
                em[b, :, c] += p
        em -= em.max(axis=2, keepdims=True)
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    sp = sparse_lists(prob) if (chain and sparse) else (None, None)
    logz, saved = pkg.hsmm.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"], d["order"], trans_pred=sp[0])
    w = rng.uniform(0.5, 1.5, size=B)
    g = torch.from_numpy(w).float().cuda()
    d_init, d_trans, d_len, d_em = pkg.hsmm.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"], g, saved, trans_succ=sp[1])
    f32 = lambda x: x.astype(np.float32).astype(np.float64)
    ref_logz, acc = O.batch_logz_and_counts(f32(prob["em"]), prob["lengths"], f32(prob["init"]), f32(prob["trans"]), f32(prob["lenp"]), prob["end"], w)
    lz = logz.cpu().numpy()
    print(shape, "pen", pen, "sparse", sparse, "logz relerr %.2e" % np.max(np.abs(lz - ref_logz) / np.abs(ref_logz)), "logz", ref_logz[:3],
          "init %.2e trans %.2e len %.2e em %.2e" % (rel_err(d_init.cpu().numpy(), acc["E_init"]), rel_err(d_trans.cpu().numpy(), acc["E_trans"]),
                                                    rel_err(d_len.cpu().numpy(), acc["E_len"]), rel_err(d_em.cpu().numpy()[:, :, :C], acc["E_em"])))
    return d_em.cpu().numpy()[:, :, :C], acc["E_em"], prob

for shape in [(9, 60, 23, 20, True, True), (9, 60, 9, 20, True, True), (5, 150, 23, 100, True, True)]:
    for pen in (0, 1e2, 1e3, 1e4):
        for sparse in (True, False):
            run(shape, pen, sparse)
mine, ref, prob = run((9, 60, 9, 20, True, True), 1e4, True)
err = np.abs(mine - ref)
b, t, c = np.unravel_index(err.argmax(), err.shape)
print("worst d_em at", b, t, c, mine[b, t, c], ref[b, t, c], "len", prob["lengths"][b])
np.set_printoptions(precision=4, suppress=True, linewidth=200)
print("mine row", mine[b, max(0,t-3):t+4])
print("ref row", ref[b, max(0,t-3):t+4])
print("em rows", prob["em"][b, max(0,t-3):t+4])
