"""Generate golden input/output vectors by running the UNMODIFIED reference HSMM module
(/root/reference/src/models/semimarkov/semimarkov_modules.py, semimarkov_utils.py) in this
container, over oracle/torch_struct_shim.py for the un-vendored pytorch-struct calls.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

The fixtures are small (a few hundred KB in total) and are committed; the GPU box only reads
the .npz files.  Everything is seeded; the reference itself seeds nothing.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_import import RefArgs, load_reference  # noqa: E402

mods, utils = load_reference()
from oracle.torch_struct_shim import MaxSemiring, SemiMarkov, SemiMarkovCRF  # noqa: E402


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


def randomise(m, gen, cov_lo=0.5):
    with torch.no_grad():
        m.gaussian_means.copy_(torch.randn(m.gaussian_means.shape, generator=gen))
        m.transition_logits.copy_(torch.randn(m.transition_logits.shape, generator=gen))
        m.init_logits.copy_(torch.randn(m.init_logits.shape, generator=gen))
        m.poisson_log_rates.copy_(torch.randn(m.poisson_log_rates.shape, generator=gen) * 0.5 + 1.0)
        D = m.gaussian_cov.shape[0]
        m.gaussian_cov.copy_(torch.diag(torch.rand(D, generator=gen) + cov_lo))


def run_case(m, feats, lengths, valid, addl_ends, constraints):
    """logZ (per video), mean-ll gradients, Viterbi spans and emissions from the reference."""
    m.zero_grad()
    vpi = None if valid is None else [valid for _ in range(feats.size(0))]
    ll, _ = m.log_likelihood(feats, lengths, vpi, spans=None, add_eos=True,
                             additional_allowed_ends_per_instance=addl_ends, constraints=constraints)
    ll.backward()
    scores, _, elp = m.score_features(feats, lengths, valid, add_eos=True, use_mean_z=True,
                                      additional_allowed_ends_per_instance=addl_ends,
                                      constraints=constraints, return_elp=True)
    logz = SemiMarkovCRF(scores, lengths + 1).partition
    spans = m.viterbi(feats, lengths, vpi, add_eos=True,
                      additional_allowed_ends_per_instance=addl_ends, constraints=constraints)
    return dict(ll=ll, logz=logz, elp=elp, viterbi_spans=spans,
                g_means=m.gaussian_means.grad, g_trans=m.transition_logits.grad,
                g_init=m.init_logits.grad, g_rates=m.poisson_log_rates.grad)


def params_of(m):
    return dict(gaussian_means=m.gaussian_means, gaussian_cov=m.gaussian_cov,
                transition_logits=m.transition_logits, init_logits=m.init_logits,
                poisson_log_rates=m.poisson_log_rates)


# ---------------------------------------------------------------------------------------------
def case_known_answer():
    """Body of models/test_semimarkov.py:266-323 (test_log_hsmm), through the reference's own
    log_hsmm; stores the decoded sequence and re-checks the reference's assertions."""
    b, C, N, K, step = 10, 4, 100, 5, 4
    BIG_NEG = -1e9
    padded = N + step * 2
    lengths_unpadded = torch.full((b,), N).long()
    lengths_unpadded[0] = padded
    lengths = lengths_unpadded + 1
    trans = torch.zeros(C, C)
    init = torch.full((C,), BIG_NEG)
    init[0] = 0
    em = torch.full((b, padded, C), BIG_NEG)
    for n in range(padded):
        em[:, n, (n // step) % C] = 1
    ls = torch.full((K, C), BIG_NEG)
    ls[step, :] = 0
    scores = mods.SemiMarkovModule.log_hsmm(trans, em, init, ls, lengths_unpadded, add_eos=True)
    marg = SemiMarkov(MaxSemiring).marginals(scores, lengths=lengths)
    seq, _ = SemiMarkov.from_parts(marg)
    for s in range(N // step):
        assert (seq[:, step * s] == s % C).all()
    assert (seq[torch.arange(b), lengths - 1] == C).all()
    save("known_answer", b=b, C=C, N=N, K=K, step=step, lengths_unpadded=lengths_unpadded, sequence=seq)


def case_labels_spans():
    """Vectors of models/test_semimarkov.py:250-263 through semimarkov_utils.py:6-63."""
    labels = torch.LongTensor([[0, 1, 1, 2, 2, 2], [0, 1, 2, 3, 3, 4]])
    spans = utils.labels_to_spans(labels, max_k=10)
    assert (spans == torch.LongTensor([[0, 1, -1, 2, -1, -1], [0, 1, 2, 3, -1, 4]])).all()
    assert (utils.spans_to_labels(spans) == labels).all()
    gen = torch.Generator().manual_seed(7)
    rand = torch.randint(0, 3, (5, 40), generator=gen)
    out = {"labels": labels, "spans_k10": spans, "rand_labels": rand}
    for k in (2, 3, 5, 50):
        out["rand_spans_k%d" % k] = utils.labels_to_spans(rand, max_k=k)
    zeros = torch.zeros(1, 6).long()
    out["zeros_k4"] = utils.labels_to_spans(zeros, max_k=4)
    save("labels_spans", **out)


def case_unconstrained():
    gen = torch.Generator().manual_seed(11)
    B, T, C, D, K = 4, 37, 5, 8, 7
    m = mods.SemiMarkovModule(RefArgs(sm_max_span_length=K), C, D, allow_self_transitions=True)
    randomise(m, gen)
    lengths = torch.LongTensor([37, 20, 9, 31])
    feats = torch.randn(B, T, D, generator=gen) + 0.3
    for i, ln in enumerate(lengths):
        feats[i, ln:] = 0  # padding_colate zero-pads (models/model.py:42-63)
    r = run_case(m, feats, lengths, None, None, None)
    save("unconstrained", features=feats, lengths=lengths, max_k=K, **params_of(m), **r)


def case_short_clamp():
    """max_k larger than the padded batch length: K is clamped (semimarkov_modules.py:450-452)."""
    gen = torch.Generator().manual_seed(13)
    B, T, C, D, K = 3, 6, 3, 4, 20
    m = mods.SemiMarkovModule(RefArgs(sm_max_span_length=K), C, D, allow_self_transitions=True)
    randomise(m, gen)
    lengths = torch.LongTensor([6, 3, 5])
    feats = torch.randn(B, T, D, generator=gen)
    for i, ln in enumerate(lengths):
        feats[i, ln:] = 0
    r = run_case(m, feats, lengths, None, None, None)
    save("short_clamp", features=feats, lengths=lengths, max_k=K, **params_of(m), **r)


def chain_module(K, D, gen):
    """Two CrossTask-like tasks with ordering chains bkg-step-bkg-... (data/crosstask.py:328-388),
    self loops (models/semimarkov/semimarkov.py:50-54) and merged backgrounds (:58-78)."""
    n_classes = 9
    chains = {0: [0, 1, 2, 3, 4], 1: [5, 6, 7]}
    allowed_starts, allowed_ends, allowed_transitions = set(), set(), {}
    for ch in chains.values():
        for s, t in zip(ch, ch[1:]):
            allowed_transitions.setdefault(s, set()).add(t)
        allowed_starts.add(ch[0])
        allowed_ends.add(ch[-1])
    for s in range(n_classes):
        allowed_transitions.setdefault(s, set()).add(s)
    merge = {0: 0, 2: 0, 4: 0, 1: 1, 3: 3, 5: 5, 7: 5, 6: 6, 8: 8}
    m = mods.SemiMarkovModule(RefArgs(sm_max_span_length=K), n_classes, D, allow_self_transitions=True,
                              allowed_starts=allowed_starts, allowed_transitions=allowed_transitions,
                              allowed_ends=allowed_ends, merge_classes=merge)
    randomise(m, gen)
    meta = dict(n_classes=n_classes, allowed_starts=sorted(allowed_starts), allowed_ends=sorted(allowed_ends),
                allowed_transitions=np.array([(s, t) for s, ts in allowed_transitions.items() for t in sorted(ts)]),
                merge_src=np.array(sorted(merge)), merge_dst=np.array([merge[k] for k in sorted(merge)]))
    return m, meta


def case_constrained(with_narration):
    gen = torch.Generator().manual_seed(17 + int(with_narration))
    B, T, D, K = 3, 26, 6, 6
    m, meta = chain_module(K, D, gen)
    valid = torch.LongTensor([0, 1, 2, 3, 4])
    lengths = torch.LongTensor([26, 4, 15])
    feats = torch.randn(B, T, D, generator=gen)
    for i, ln in enumerate(lengths):
        feats[i, ln:] = 0
    # video 1 is shorter than the chain: extra allowed end (semimarkov.py:135-147)
    addl = [[], [3], []]
    constraints = None
    if with_narration:
        # (1 - c) * -1e4 scattered into the step columns (semimarkov.py:149-157, 227-232)
        c = torch.zeros(B, T, len(valid))
        for i in range(B):
            for col, (lo, hi) in zip((1, 3), ((2, 12), (10, 24))):
                allowed = torch.zeros(T)
                allowed[lo:hi] = 1
                c[i, :, col] = (1 - allowed) * -1e4
        constraints = c
    r = run_case(m, feats, lengths, valid, addl, constraints)
    extra = {} if constraints is None else {"constraints": constraints}
    save("constrained_narration" if with_narration else "constrained", features=feats, lengths=lengths,
         max_k=K, valid_classes=valid, addl_ends_flat=np.array([-1, 3, -1]),
         init_constraints=m.init_constraints, transition_constraints=m.transition_constraints,
         **meta, **params_of(m), **r, **extra)


def case_supervised_fit():
    """fit_supervised closed form (semimarkov_modules.py:195-256, semimarkov_utils.py:74-126)."""
    gen = torch.Generator().manual_seed(23)
    C, D, K, n_vid = 4, 6, 8, 12
    m = mods.SemiMarkovModule(RefArgs(sm_max_span_length=K), C, D, allow_self_transitions=True)
    feats, labels, lens = [], [], []
    mu = torch.randn(C, D, generator=gen) * 2
    for i in range(n_vid):
        T = int(torch.randint(10, 40, (1,), generator=gen))
        lab, cur = [], int(torch.randint(0, C, (1,), generator=gen))
        while len(lab) < T:
            lab.extend([cur] * int(torch.randint(1, 12, (1,), generator=gen)))
            cur = (cur + int(torch.randint(1, C, (1,), generator=gen))) % C
        lab = torch.LongTensor(lab[:T])
        feats.append(mu[lab] + torch.randn(T, D, generator=gen))
        labels.append(lab)
        lens.append(T)
    m.fit_supervised(feats, labels)
    save("supervised_fit", features=torch.cat(feats), labels=torch.cat(labels), lengths=np.array(lens),
         max_k=K, n_classes=C, **params_of(m))
    # decode the training videos with the fitted model (S6 test-time path)
    Tm = max(lens)
    X = torch.zeros(n_vid, Tm, D)
    for i, f in enumerate(feats):
        X[i, :lens[i]] = f
    spans = m.viterbi(X, torch.LongTensor(lens), None, add_eos=True)
    save("supervised_decode", features=X, lengths=np.array(lens), max_k=K, viterbi_spans=spans, **params_of(m))


def case_gold_score():
    """log_likelihood with gold spans: generative score and discriminative log-prob
    (semimarkov_modules.py:626-655), equal-length batch (ragged gold batches read padding)."""
    gen = torch.Generator().manual_seed(29)
    B, T, C, D, K = 3, 18, 4, 5, 6
    out = {}
    for disc in (False, True):
        m = mods.SemiMarkovModule(RefArgs(sm_max_span_length=K, sm_train_discriminatively=disc), C, D,
                                  allow_self_transitions=True)
        randomise(m, torch.Generator().manual_seed(31))
        lengths = torch.LongTensor([T] * B)
        feats = torch.randn(B, T, D, generator=torch.Generator().manual_seed(37))
        labels = torch.randint(0, C, (B, T), generator=torch.Generator().manual_seed(41))
        labels = labels.sort(dim=1)[0]
        spans = utils.labels_to_spans(labels, max_k=K)
        m.zero_grad()
        ll, _ = m.log_likelihood(feats, lengths, None, spans=spans, add_eos=True)
        ll.backward()
        tag = "disc" if disc else "gen"
        out.update({"ll_" + tag: ll, "g_means_" + tag: m.gaussian_means.grad, "g_trans_" + tag: m.transition_logits.grad,
                    "g_init_" + tag: m.init_logits.grad, "g_rates_" + tag: m.poisson_log_rates.grad})
        if not disc:
            out.update(features=feats, lengths=lengths, labels=labels, spans=spans, max_k=K, **params_of(m))
    save("gold_score", **out)


if __name__ == "__main__":
    case_known_answer()
    case_labels_spans()
    case_unconstrained()
    case_short_clamp()
    case_constrained(False)
    case_constrained(True)
    case_supervised_fit()
    case_gold_score()
