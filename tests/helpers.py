"""Shared helpers for the GPU parity tests (oracle = checker, product = CUDA path)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import hsmm_oracle as O  # noqa: E402
from action_segmentation_b200.args import HsmmArgs as RefArgs  # noqa: E402


def module_from_golden(g, device="cuda", **argkw):
    import action_segmentation_b200 as pkg
    n_classes = int(g["init_logits"].shape[0])
    D = int(g["gaussian_means"].shape[1])
    kw = {}
    if "allowed_starts" in g:
        trans = {}
        for s, t in g["allowed_transitions"]:
            trans.setdefault(int(s), set()).add(int(t))
        kw = dict(allowed_starts=set(int(x) for x in g["allowed_starts"]), allowed_transitions=trans,
                  allowed_ends=set(int(x) for x in g["allowed_ends"]),
                  merge_classes={int(s): int(d) for s, d in zip(g["merge_src"], g["merge_dst"])})
    m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=int(g["max_k"]), **argkw), n_classes, D,
                             allow_self_transitions=True, **kw)
    with torch.no_grad():
        for k in ("gaussian_means", "gaussian_cov", "transition_logits", "init_logits", "poisson_log_rates"):
            getattr(m, k).copy_(torch.from_numpy(g[k]))
    return m.to(device)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def random_problem(rng, B, Tmax, C, K, Tmin=1, chain=False, ends=False, scale=3.0, narration=False):
    """Random DP inputs (already in score space): em (B,Tmax,C), init, trans [to,from], lenp (K,C), end."""
    lengths = rng.integers(Tmin, Tmax + 1, size=B)
    lengths[0] = Tmax
    em = rng.normal(size=(B, Tmax, C)) * scale
    if narration:
        # soft -1e4 penalty on the odd ("step") classes outside one window each (semimarkov.py:227-232)
        for b in range(B):
            for c in range(1, C, 2):
                lo = int(rng.integers(0, max(1, lengths[b])))
                hi = lo + int(rng.integers(2, 12))
                pen = np.full(Tmax, -1e4)
                pen[lo:hi] = 0.0
                em[b, :, c] += pen
    em -= em.max(axis=2, keepdims=True)
    init = np.log(rng.dirichlet(np.ones(C)))
    logits = rng.normal(size=(C, C))
    if chain:
        mask = np.full((C, C), True)
        for c in range(C):
            mask[c, c] = False
            if c + 1 < C:
                mask[c + 1, c] = False
        logits = np.where(mask, O.BIG_NEG, logits)
        init = O.log_softmax(np.where(np.arange(C) == 0, 0.0, O.BIG_NEG), axis=0)
    trans = O.log_softmax(logits, axis=0)
    lenp = O.poisson_length_log_probs(np.log(rng.uniform(1.0, max(2.0, K / 2.0), size=C)), K)
    end = None
    if ends:
        end = np.full((B, C), O.BIG_NEG)
        end[:, C - 1] = 0.0
        for b in range(B):
            if lengths[b] < C:
                end[b, int(lengths[b]) - 1] = 0.0
    return dict(em=em, lengths=lengths, init=init, trans=trans, lenp=lenp, end=end)


def sparse_lists(prob, device="cuda"):
    """(pred, succ) hint tensors from the unmasked (> -1e8) entries of prob['trans']."""
    import action_segmentation_b200 as pkg
    allowed = torch.from_numpy(prob["trans"] > -1e8)
    return pkg.hsmm.sparse_transition_lists(allowed, torch.device(device))


def to_dev(prob, device="cuda"):
    """Device tensors for the low-level entry points (em padded to ldc)."""
    import action_segmentation_b200 as pkg
    B, T, C = prob["em"].shape
    ldc = pkg.hsmm.ldc_of(C)
    em = torch.zeros(B, T, ldc, device=device)
    em[:, :, :C] = torch.from_numpy(prob["em"]).float()
    f = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x)).float().to(device)  # noqa: E731
    lengths = torch.from_numpy(prob["lengths"]).long()
    lengths_i32, order = pkg.hsmm.prepare_lengths(lengths, torch.device(device))
    return dict(em=em, init=f(prob["init"]), trans=f(prob["trans"]), lenp=f(prob["lenp"]), end=f(prob["end"]),
                lengths_i32=lengths_i32, order=order, C=C)


def check_viterbi_against_oracle(prob, spans, score=None, tol=1e-4, em_round=True, stats=None):
    """Exact path match unless the CUDA path is a numerical near-tie: its fp64 score must be within
    tol (relative) of the oracle's best score.  `spans` (B,Tmax+1) in local ids, EOS = C."""
    em = prob["em"].astype(np.float32).astype(np.float64) if em_round else prob["em"]
    B, Tmax, C = em.shape
    lenp = O.clamp_len_table(prob["lenp"].astype(np.float32).astype(np.float64), Tmax)
    init = prob["init"].astype(np.float32).astype(np.float64)
    trans = prob["trans"].astype(np.float32).astype(np.float64)
    n_exact = 0
    same_frames = total_frames = 0
    for b in range(B):
        T = int(prob["lengths"][b])
        end = None if prob["end"] is None else prob["end"][b]
        best, segs = O.viterbi(em[b, :T], init, trans, lenp, end)
        ref_row = O.segs_to_spans(segs, T, C, Tmax + 1)
        row = np.asarray(spans[b])
        assert row[T] == C, "EOS marker missing"
        assert (row[T + 1:] == -1).all(), "padding positions must be -1"
        if (row == ref_row).all():
            n_exact += 1
        same_frames += int((O.spans_to_labels(row[None, :T])[0] == O.spans_to_labels(ref_row[None, :T])[0]).sum())
        total_frames += T
        mine = O.segments_from_spans(row, T)
        assert all(1 <= ln <= lenp.shape[0] - 1 for _, ln, _ in mine), "segment longer than K-1"
        s = O.path_score(mine, em[b, :T], init, trans, lenp, end)
        assert best - s <= tol * max(1.0, abs(best)), (b, best, s)
        if score is not None:
            assert abs(float(score[b]) - s) <= 1e-5 * max(1.0, abs(s)) + 1e-3, (b, float(score[b]), s)
    if stats is not None:
        stats["frame_agreement"] = same_frames / max(1, total_frames)
    return n_exact
