"""Latency / saturation sweep of the DP kernels: time per launch vs number of videos, fixed-length videos.
Usage: python tools/perf_sweep.py [C K T]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import action_segmentation_b200 as pkg  # noqa: E402
from tests.helpers import random_problem, sparse_lists, to_dev  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    C, K, T = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (23, 20, 2000)
    chain = (len(sys.argv) < 5) or sys.argv[4] != "dense"
    rng = np.random.default_rng(0)
    for V in (128, 592, 2368, 9472):
        prob = random_problem(rng, V, T, C, K, Tmin=T, chain=chain, ends=chain)
        d = to_dev(prob)
        sp = sparse_lists(prob) if chain else (None, None)
        g = torch.ones(V, device="cuda")
        H = pkg.hsmm
        st = {}
        t_f = timed(lambda: st.__setitem__("f", H.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None,
                                                               d["lengths_i32"], d["order"], trans_pred=sp[0])))
        logz, saved = st["f"]
        t_b = timed(lambda: H.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"],
                                            g, saved, trans_succ=sp[1]))
        t_v = timed(lambda: H.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"],
                                             d["order"], want_score=False, trans_pred=sp[0]))
        fr = V * T
        clk = 1.965e9
        print("C=%d K=%d T=%d V=%5d | fwd %8.3f ms (%6.0f clk/frame-step, %7.1f Mfr/s) | bwd %8.3f ms (%6.0f clk, %7.1f Mfr/s) | "
              "vit %8.3f ms (%6.0f clk, %7.1f Mfr/s) | %s" % (
                  C, K, T, V, t_f, t_f * 1e-3 * clk / T, fr / t_f / 1e3, t_b, t_b * 1e-3 * clk / T, fr / t_b / 1e3,
                  t_v, t_v * 1e-3 * clk / T, fr / t_v / 1e3, pkg._lib.dp_variant(C, K, 1, chain)), flush=True)


if __name__ == "__main__":
    main()
