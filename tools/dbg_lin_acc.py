"""Accuracy of the linear-window and log-domain kernels against the fp64 oracle on one LIN_SHAPES case."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import action_segmentation_b200 as pkg  # noqa: E402
from oracle import hsmm_oracle as O  # noqa: E402
from tests.helpers import random_problem, rel_err, sparse_lists, to_dev  # noqa: E402
from tests.test_gpu_parity import LIN_SHAPES, _fwd_bwd  # noqa: E402

for shape in LIN_SHAPES[:4]:
    B, Tmax, C, K, chain, ends, scale = shape
    rng = np.random.default_rng(300 + C * 7 + K)
    prob = random_problem(rng, B, Tmax, C, K, Tmin=Tmax // 3, chain=chain, ends=ends, scale=scale)
    prob["lenp"] = O.poisson_length_log_probs(np.log(rng.uniform(1.0, 12.0, size=C)), K)
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    sp = sparse_lists(prob) if chain else (None, None)
    w = rng.uniform(0.5, 1.5, size=B)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    ref_logz, acc = O.batch_logz_and_counts(f32(prob["em"]), prob["lengths"], f32(prob["init"]), f32(prob["trans"]),
                                            f32(prob["lenp"]), prob["end"], w)
    for mode in (True, False):
        pkg._lib.set_linear_window(mode)
        lz, g, ff, bf = _fwd_bwd(prob, d, sp, w)
        mine = dict(E_init=g[0], E_trans=g[1], E_len=g[2], E_em=g[3][:, :, :C])
        errs = {k: rel_err(v.cpu().numpy(), acc[k]) for k, v in mine.items()}
        e_em = np.abs(g[3][:, :, :C].cpu().numpy() - acc["E_em"])
        bad = np.unravel_index(e_em.argmax(), e_em.shape)
        print(shape, "lin" if mode else "log", "logz err %.2e" % np.abs(lz.cpu().numpy() - ref_logz).max(),
              {k: "%.1e" % v for k, v in errs.items()}, "worst E_em at", bad, "flags", ff.max(), bf.max())
    pkg._lib.set_linear_window(True)
