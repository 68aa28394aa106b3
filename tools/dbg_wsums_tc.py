"""Debug aid for the tensor-core weighted-sums kernel: compares against numpy on a few shapes and prints where the
error sits (feature chunk x class), and A/B timing against the SIMT kernel via HSMM_WSUMS_TC."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import action_segmentation_b200 as pkg  # noqa: E402


def run(B, Tmax, D, C, seed=0, full=False):
    rng = np.random.default_rng(seed)
    lengths = np.full(B, Tmax) if full else rng.integers(1, Tmax + 1, size=B)
    lengths[0] = Tmax
    X = rng.normal(size=(B, Tmax, D)).astype(np.float32)
    ldc = pkg.hsmm.ldc_of(C)
    wgt = np.zeros((B, Tmax, ldc), dtype=np.float32)
    wgt[:, :, :C] = rng.dirichlet(np.ones(C), size=(B, Tmax))
    li = torch.from_numpy(lengths).to(torch.int32).cuda()
    wx, wsum = pkg.hsmm.weighted_feature_sums(torch.from_numpy(X).cuda(), torch.from_numpy(wgt).cuda(), C, li)
    torch.cuda.synchronize()
    ref_wx = np.zeros((C, D))
    ref_ws = np.zeros(C)
    for b, T in enumerate(lengths):
        ref_wx += wgt[b, :T, :C].astype(np.float64).T @ X[b, :T].astype(np.float64)
        ref_ws += wgt[b, :T, :C].astype(np.float64).sum(axis=0)
    err = np.abs(wx.cpu().numpy() - ref_wx)
    scale = np.sqrt(float(lengths.sum()))
    print("B=%d T=%d D=%d C=%d: max err %.3g (tol %.3g), wsum rel err %.3g" % (
        B, Tmax, D, C, err.max(), 2e-5 * scale, np.abs(wsum.cpu().numpy() - ref_ws).max() / np.abs(ref_ws).max()))
    if err.max() > 2e-5 * scale:
        nch = (D + 31) // 32
        tab = np.array([[err[c, ch * 32:(ch + 1) * 32].max() for ch in range(nch)] for c in range(C)])
        np.set_printoptions(precision=2, linewidth=200, suppress=False)
        print("err by class (rows) x feature chunk (cols):")
        print(tab)
        print("mine[0,:8]", wx.cpu().numpy()[0, :8], "\nref [0,:8]", ref_wx[0, :8])


if __name__ == "__main__":
    run(1, 32, 200, 23, full=True)
    if len(sys.argv) > 1:
        run(1, 8, 32, 4, full=True)
        run(3, 100, 200, 23)
        run(6, 333, 200, 23)
        run(5, 200, 64, 7)
        run(4, 77, 8, 5)
