"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

numpy (fp64 by default) restatement of the reference's HSMM hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.

Every function cites the reference lines it follows (paths relative to /root/reference/src).
PARITY STATUS
  * scoring functions (emissions, Poisson lengths, masked log-softmax, span encodings,
    supervised statistics) are pinned against the reference's own code run in this
    container: tests/golden/make_golden.py imports models/semimarkov/semimarkov_modules.py
    unmodified and the resulting fixtures are checked by tests/test_oracle_golden.py.
  * the DP (pytorch-struct, un-vendored third party) is "parity unpinned" against the real
    library; it is pinned by the reference's known-answer test (models/test_semimarkov.py:266-323)
    and by brute-force enumeration (``brute_force`` below).

The mathematical object:

    score(segmentation) = init[c_0] + sum_i ( len[l_i, c_i] + sum_{t in seg_i} em[t, c_i] )
                        + sum_{i>=1} trans[c_i, c_{i-1}] + end[c_last]

segments tile [0, T) exactly, l_i in 1..K-1, trans indexed [to, from].
"""
import itertools
import math

import numpy as np

BIG_NEG = -1e9  # models/semimarkov/semimarkov_modules.py:20
LOG_2PI = math.log(2.0 * math.pi)


# ------------------------------------------------------------------------------------------
# small numerics helpers
# ------------------------------------------------------------------------------------------
def logsumexp(x, axis=None):
    x = np.asarray(x)
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    out = np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m
    if axis is None:
        return out.reshape(())
    return np.squeeze(out, axis=axis)


def log_softmax(x, axis=0):
    return x - np.expand_dims(logsumexp(x, axis=axis), axis)


# ------------------------------------------------------------------------------------------
# parameter -> score transforms
# ------------------------------------------------------------------------------------------
def initial_log_probs(init_logits, init_constraints=None, valid_classes=None):
    """models/semimarkov/semimarkov_modules.py:284-296."""
    logits = np.array(init_logits, dtype=np.float64)
    if init_constraints is not None:
        logits = np.where(np.asarray(init_constraints, dtype=bool), BIG_NEG, logits)
    if valid_classes is not None:
        logits = logits[np.asarray(valid_classes)]
    return log_softmax(logits, axis=0)


def transition_log_probs(transition_logits, transition_constraints=None, valid_classes=None,
                         allow_self_transitions=True):
    """models/semimarkov/semimarkov_modules.py:298-322.  Indexed [to, from]; columns normalised."""
    logits = np.array(transition_logits, dtype=np.float64)
    if transition_constraints is not None:
        logits = np.where(np.asarray(transition_constraints, dtype=bool), BIG_NEG, logits)
    if valid_classes is not None:
        vc = np.asarray(valid_classes)
        logits = logits[vc][:, vc]
    if not allow_self_transitions:
        logits = np.where(np.eye(logits.shape[0], dtype=bool), BIG_NEG, logits)
    return log_softmax(logits, axis=0)


def poisson_length_log_probs(log_rates, max_k):
    """models/semimarkov/semimarkov_modules.py:383-398: rows k = 0..max_k-1, not renormalised."""
    log_rates = np.asarray(log_rates, dtype=np.float64)
    if max_k == 1:
        return np.tile(np.array([[0.0], [-1000.0]]), (1, log_rates.shape[-1]))
    k = np.arange(max_k, dtype=np.float64)[:, None]
    lgam = np.array([math.lgamma(i + 1.0) for i in range(max_k)])[:, None]
    return k * log_rates[None, :] - np.exp(log_rates)[None, :] - lgam


def emission_log_probs(features, class_means, cov_diag, constraints=None):
    """models/semimarkov/semimarkov_modules.py:324-381 with a tied diagonal covariance:
    log N(x; mu_c, diag(var)) (+ additive constraints).  features (..., T, D), means (C, D)."""
    x = np.asarray(features, dtype=np.float64)
    mu = np.asarray(class_means, dtype=np.float64)
    var = np.asarray(cov_diag, dtype=np.float64)
    D = x.shape[-1]
    diff = x[..., None, :] - mu  # (..., T, C, D)
    maha = np.sum(diff * diff / var, axis=-1)
    elp = -0.5 * maha - 0.5 * np.sum(np.log(var)) - 0.5 * D * LOG_2PI
    if constraints is not None:
        elp = elp + np.asarray(constraints, dtype=np.float64)
    return elp


# ------------------------------------------------------------------------------------------
# span encodings  (models/semimarkov/semimarkov_utils.py:6-63)
# ------------------------------------------------------------------------------------------
def labels_to_spans(labels, max_k):
    """semimarkov_utils.py:6-23: class id at segment starts, -1 inside; runs split at max_k-1."""
    labels = np.asarray(labels)
    assert not (labels == -1).any()
    out = np.empty_like(labels)
    B, N = labels.shape
    for b in range(B):
        run = 0
        for n in range(N):
            same = n > 0 and labels[b, n] == labels[b, n - 1]
            if max_k is not None:
                same = same and run < max_k - 1
            if same:
                out[b, n] = -1
                run += 1
            else:
                out[b, n] = labels[b, n]
                run = 1
    return out


def spans_to_labels(spans):
    """semimarkov_utils.py:51-63."""
    spans = np.asarray(spans)
    out = np.empty_like(spans)
    assert (spans[:, 0] != -1).all()
    cur = spans[:, 0].copy()
    for n in range(spans.shape[1]):
        cur = np.where(spans[:, n] == -1, cur, spans[:, n])
        out[:, n] = cur
    return out


def rle_spans(spans, lengths):
    """semimarkov_utils.py:26-48."""
    out = []
    for row, ln in zip(np.asarray(spans), lengths):
        rle = []
        for sym in row[:int(ln)]:
            sym = int(sym)
            if not rle or sym != -1:
                rle.append([sym, 0])
            rle[-1][1] += 1
        out.append([tuple(x) for x in rle])
    return out


def segments_from_spans(span_row, length):
    """[(start, length, class)] of one span-encoded row (first `length` positions)."""
    segs = []
    for t in range(int(length)):
        s = int(span_row[t])
        if s != -1:
            segs.append([t, 0, s])
        segs[-1][1] += 1
    return [tuple(s) for s in segs]


# ------------------------------------------------------------------------------------------
# dense potentials + the materialised DP (the reference's own formulation; small sizes only)
# ------------------------------------------------------------------------------------------
def sliding_sum(x, k):
    """models/semimarkov/semimarkov_modules.py:26-39: out[i] = sum_{t=i}^{i+k-1} x[t], zero padded."""
    T = x.shape[0]
    pad = np.concatenate([x, np.zeros((k,) + x.shape[1:], dtype=x.dtype)], axis=0)
    cs = np.concatenate([np.zeros((1,) + x.shape[1:], dtype=x.dtype), np.cumsum(pad, axis=0)], axis=0)
    return cs[k:k + T] - cs[:T]


def log_hsmm(transition, emission, init, length_scores, lengths, allowed_ends_per_instance=None):
    """models/semimarkov/semimarkov_modules.py:416-523 with add_eos=True, all_batched=False.
    Returns scores (B, Tmax, K, C+1, C+1)."""
    em = np.asarray(emission, dtype=np.float64)
    B, N1, C1 = em.shape
    K = length_scores.shape[0]
    if K > N1:
        K = N1
        length_scores = length_scores[:K]
    N, C = N1 + 1, C1 + 1
    tr = np.full((B, C, C), BIG_NEG)
    tr[:, :C1, :C1] = transition
    if allowed_ends_per_instance is None:
        tr[:, C1, :] = 0
    else:
        for i, ends in enumerate(allowed_ends_per_instance):
            tr[i, C1, list(ends)] = 0
    ini = np.full((B, C), BIG_NEG)
    ini[:, :C1] = init
    ls = np.full((B, K, C), BIG_NEG)
    ls[:, :, :C1] = length_scores
    if K > 1:
        ls[:, 1, C1] = 0
    else:
        ls[:, 0, C1] = 0
    ea = np.full((B, N, C), BIG_NEG)
    for i, ln in enumerate(lengths):
        ea[i, :ln, :C1] = em[i, :ln]
        ea[i, ln, C1] = 0
    scores = np.zeros((B, N - 1, K, C, C))
    scores += tr[:, None, None]
    scores[:, 0] += ini[:, None, None, :]
    scores += ls[:, None, :, None, :]
    for k in range(1, K):
        for i in range(B):
            ln = int(lengths[i]) + 1
            summed = sliding_sum(ea[i], k)  # (N, C)
            scores[i, :ln - 1, k] += summed[:ln - 1][:, None, :]
            scores[i, ln - 1 - k, k] += ea[i, ln - 1][:, None]  # may wrap (negative index) as in the reference
    return scores


def materialised_dp(scores, lengths_with_eos, semiring="log"):
    """pytorch-struct SemiMarkov._dp restated (oracle/torch_struct_shim.py docstring)."""
    B, N1, K, C, _ = scores.shape
    out = np.zeros(B)
    for b in range(B):
        ln = int(lengths_with_eos[b])
        beta = [np.zeros(C)]
        alpha = []
        for n in range(1, ln):
            e = scores[b, n - 1] + beta[n - 1][None, None, :]  # (K, C2, C1)
            alpha.append(logsumexp(e, axis=-1) if semiring == "log" else e.max(axis=-1))
            stack = np.stack([alpha[n - k][k] for k in range(1, min(K - 1, n) + 1)], axis=0)
            beta.append(logsumexp(stack, axis=0) if semiring == "log" else stack.max(axis=0))
        out[b] = logsumexp(beta[ln - 1]) if semiring == "log" else beta[ln - 1].max()
    return out


# ------------------------------------------------------------------------------------------
# factorised O(T C (L + C)) forward / backward / Viterbi for ONE video
# ------------------------------------------------------------------------------------------
def _prep(em, init, trans, lenp, end, dtype):
    em = np.asarray(em, dtype=dtype)
    T, C = em.shape
    init = np.asarray(init, dtype=dtype)
    trans = np.asarray(trans, dtype=dtype)
    lenp = np.asarray(lenp, dtype=dtype)
    end = np.zeros(C, dtype=dtype) if end is None else np.asarray(end, dtype=dtype)
    L = lenp.shape[0] - 1
    P = np.concatenate([np.zeros((1, C), dtype=dtype), np.cumsum(em, axis=0, dtype=dtype)], axis=0)
    return em, init, trans, lenp, end, T, C, L, P


def forward(em, init, trans, lenp, end=None, dtype=np.float64):
    """SURVEY.md 3.5 factorised forward.  lenp has rows k = 0..K-1 (row 0 unused).
    Returns logZ, beta (T, C) [beta[n] for n=0..T-1], gamma (T+1, C) [gamma[0] unused]."""
    em, init, trans, lenp, end, T, C, L, P = _prep(em, init, trans, lenp, end, dtype)
    beta = np.full((T, C), -np.inf, dtype=dtype)
    gamma = np.full((T + 1, C), -np.inf, dtype=dtype)
    beta[0] = init
    for n in range(1, T + 1):
        kmax = min(L, n)
        ks = np.arange(1, kmax + 1)
        cand = beta[n - ks] + lenp[ks] + (P[n][None, :] - P[n - ks])
        gamma[n] = logsumexp(cand, axis=0)
        if n < T:
            beta[n] = logsumexp(gamma[n][None, :] + trans, axis=1)
    logZ = logsumexp(gamma[T] + end)
    return logZ, beta, gamma


def backward(em, init, trans, lenp, end=None, dtype=np.float64):
    """Backward recursion: eta[n][c] completion after a class-c segment ended at n,
    zeta[n][c] completion when a class-c segment starts at n."""
    em, init, trans, lenp, end, T, C, L, P = _prep(em, init, trans, lenp, end, dtype)
    eta = np.full((T + 1, C), -np.inf, dtype=dtype)
    zeta = np.full((T, C), -np.inf, dtype=dtype)
    eta[T] = end
    for n in range(T - 1, -1, -1):
        kmax = min(L, T - n)
        ks = np.arange(1, kmax + 1)
        cand = lenp[ks] + (P[n + ks] - P[n][None, :]) + eta[n + ks]
        zeta[n] = logsumexp(cand, axis=0)
        if n > 0:
            eta[n] = logsumexp(trans + zeta[n][:, None], axis=0)
    return eta, zeta


def expected_counts(em, init, trans, lenp, end=None, dtype=np.float64):
    """logZ and d logZ / d {init, trans, len, em, end}: the marginals the reference obtains by
    autograd through pytorch-struct + log_hsmm (semimarkov.py:284-286)."""
    em, init, trans, lenp, end, T, C, L, P = _prep(em, init, trans, lenp, end, dtype)
    logZ, beta, gamma = forward(em, init, trans, lenp, end, dtype)
    eta, zeta = backward(em, init, trans, lenp, end, dtype)
    E_len = np.zeros_like(lenp)
    occ_delta = np.zeros((T + 1, C), dtype=dtype)
    for n in range(T):
        kmax = min(L, T - n)
        ks = np.arange(1, kmax + 1)
        q = np.exp(beta[n][None, :] + lenp[ks] + (P[n + ks] - P[n][None, :]) + eta[n + ks] - logZ)
        E_len[ks] += q
        occ_delta[n] += q.sum(axis=0)
        np.subtract.at(occ_delta, n + ks, q)
    E_em = np.cumsum(occ_delta, axis=0)[:T]
    E_init = np.exp(init + zeta[0] - logZ)
    E_trans = np.zeros((C, C), dtype=dtype)
    for n in range(1, T):
        E_trans += np.exp(gamma[n][None, :] + trans + zeta[n][:, None] - logZ)
    E_end = np.exp(gamma[T] + end - logZ)
    return dict(logZ=logZ, E_init=E_init, E_trans=E_trans, E_len=E_len, E_em=E_em, E_end=E_end,
                beta=beta, gamma=gamma, eta=eta, zeta=zeta)


def viterbi(em, init, trans, lenp, end=None, dtype=np.float64):
    """Max-plus version with back-pointers.  Ties: smallest k, then smallest c1.
    Returns (best score, [(start, length, class), ...])."""
    em, init, trans, lenp, end, T, C, L, P = _prep(em, init, trans, lenp, end, dtype)
    beta = np.full((T, C), -np.inf, dtype=dtype)
    gamma = np.full((T + 1, C), -np.inf, dtype=dtype)
    bpk = np.zeros((T + 1, C), dtype=np.int64)
    bpc = np.zeros((T, C), dtype=np.int64)
    beta[0] = init
    for n in range(1, T + 1):
        kmax = min(L, n)
        ks = np.arange(1, kmax + 1)
        cand = beta[n - ks] + lenp[ks] + (P[n][None, :] - P[n - ks])
        a = cand.argmax(axis=0)
        bpk[n] = ks[a]
        gamma[n] = cand[a, np.arange(C)]
        if n < T:
            m = gamma[n][None, :] + trans
            bpc[n] = m.argmax(axis=1)
            beta[n] = m.max(axis=1)
    fin = gamma[T] + end
    c = int(fin.argmax())
    best = fin[c]
    segs = []
    n = T
    while n > 0:
        k = int(bpk[n, c])
        segs.append((n - k, k, c))
        n -= k
        if n > 0:
            c = int(bpc[n, c])
    return best, segs[::-1]


def path_score(segs, em, init, trans, lenp, end=None, dtype=np.float64):
    """Score of one explicit segmentation [(start, length, class)]."""
    em, init, trans, lenp, end, T, C, L, P = _prep(em, init, trans, lenp, end, dtype)
    s = init[segs[0][2]]
    pos = 0
    prev = None
    for (st, ln, c) in segs:
        assert st == pos and 1 <= ln <= L, (st, pos, ln, L)
        s += lenp[ln, c] + (P[st + ln, c] - P[st, c])
        if prev is not None:
            s += trans[c, prev]
        prev = c
        pos += ln
    assert pos == T
    return s + end[prev]


def segs_to_spans(segs, T, eos_id, total_len=None):
    """Reference span encoding of a decoded video: class at starts, -1 inside, EOS at T."""
    total_len = T + 1 if total_len is None else total_len
    row = np.full(total_len, -1, dtype=np.int64)
    for (st, ln, c) in segs:
        row[st] = c
    row[T] = eos_id
    return row


def brute_force(em, init, trans, lenp, end=None):
    """Enumerate every segmentation (tiny T only).  Returns logZ, best score, best segs and the
    dict of marginals computed from explicit path probabilities."""
    em, init, trans, lenp, end, T, C, L, P = _prep(em, init, trans, lenp, end, np.float64)

    def compositions(total):
        if total == 0:
            yield []
            return
        for first in range(1, min(L, total) + 1):
            for rest in compositions(total - first):
                yield [first] + rest

    scores, paths = [], []
    for comp in compositions(T):
        for labels in itertools.product(range(C), repeat=len(comp)):
            segs, pos = [], 0
            for ln, c in zip(comp, labels):
                segs.append((pos, ln, c))
                pos += ln
            scores.append(path_score(segs, em, init, trans, lenp, end))
            paths.append(segs)
    scores = np.array(scores)
    logZ = logsumexp(scores)
    post = np.exp(scores - logZ)
    E_init = np.zeros(C)
    E_trans = np.zeros((C, C))
    E_len = np.zeros_like(lenp)
    E_em = np.zeros((T, C))
    for p, segs in zip(post, paths):
        E_init[segs[0][2]] += p
        prev = None
        for (st, ln, c) in segs:
            E_len[ln, c] += p
            E_em[st:st + ln, c] += p
            if prev is not None:
                E_trans[c, prev] += p
            prev = c
    ibest = int(scores.argmax())
    srt = np.sort(scores)
    gap = srt[-1] - srt[-2] if len(srt) > 1 else np.inf
    return dict(logZ=logZ, best=scores[ibest], best_segs=paths[ibest], gap=gap,
                E_init=E_init, E_trans=E_trans, E_len=E_len, E_em=E_em)


# ------------------------------------------------------------------------------------------
# batch-level convenience mirroring the reference call signatures
# ------------------------------------------------------------------------------------------
def clamp_len_table(lenp, t_max):
    """`if K > N_1: K = N_1` (models/semimarkov/semimarkov_modules.py:450-452)."""
    K = lenp.shape[0]
    return lenp[:min(K, t_max)]


def end_scores(C, allowed_ends=None):
    """EOS transition row of log_hsmm (semimarkov_modules.py:462-471): 0 if class may end else -1e9."""
    if allowed_ends is None:
        return np.zeros(C)
    e = np.full(C, BIG_NEG)
    e[list(allowed_ends)] = 0.0
    return e


def batch_logz_and_counts(em, lengths, init, trans, lenp, ends=None, weights=None, dtype=np.float64):
    """Per-video logZ plus weight-summed expected counts over a ragged batch.
    em (B, Tmax, C); ends (B, C) or None; weights (B,) multiply each video's counts."""
    B, Tmax, C = em.shape
    lenp = clamp_len_table(np.asarray(lenp), Tmax)
    logz = np.zeros(B)
    acc = dict(E_init=np.zeros(C), E_trans=np.zeros((C, C)), E_len=np.zeros(lenp.shape),
               E_em=np.zeros((B, Tmax, C)))
    for b in range(B):
        T = int(lengths[b])
        r = expected_counts(em[b, :T], init, trans, lenp, None if ends is None else ends[b], dtype)
        w = 1.0 if weights is None else float(weights[b])
        logz[b] = r["logZ"]
        acc["E_init"] += w * r["E_init"]
        acc["E_trans"] += w * r["E_trans"]
        acc["E_len"] += w * r["E_len"]
        acc["E_em"][b, :T] = w * r["E_em"]
    return logz, acc


def batch_viterbi(em, lengths, init, trans, lenp, ends=None, dtype=np.float64):
    """Span-encoded predictions (B, Tmax+1) in local ids (EOS = C) and per-video best scores."""
    B, Tmax, C = em.shape
    lenp = clamp_len_table(np.asarray(lenp), Tmax)
    spans = np.full((B, Tmax + 1), -1, dtype=np.int64)
    best = np.zeros(B)
    for b in range(B):
        T = int(lengths[b])
        best[b], segs = viterbi(em[b, :T], init, trans, lenp, None if ends is None else ends[b], dtype)
        spans[b] = segs_to_spans(segs, T, C, Tmax + 1)
    return spans, best
