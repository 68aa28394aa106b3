"""Pin the oracle (oracle/hsmm_oracle.py, oracle/module_oracle.py) against the golden vectors that
tests/golden/make_golden.py produced by running the unmodified reference module.  CPU only."""
import numpy as np
import pytest

from oracle import hsmm_oracle as O
from oracle.module_oracle import from_golden, golden_addl_ends

CASES = ["unconstrained", "short_clamp", "constrained", "constrained_narration"]
KEYMAP = dict(g_means="gaussian_means", g_trans="transition_logits", g_init="init_logits", g_rates="poisson_log_rates")


def rel(a, b):
    return np.abs(a - b).max() / max(1e-12, np.abs(b).max())


@pytest.mark.parametrize("case", CASES)
def test_loglik_and_grads(golden, case):
    g = golden(case)
    mo = from_golden(g)
    r = mo.log_likelihood(g["features"], g["lengths"], g.get("valid_classes"), golden_addl_ends(g), g.get("constraints"))
    # reference ran in fp32; elp magnitudes reach 1e4 under narration penalties
    assert rel(r["elp"], g["elp"]) < 2e-6
    assert np.allclose(r["logz"], g["logz"], rtol=2e-6, atol=1e-4)
    assert abs(r["ll"] - float(g["ll"])) < 2e-6 * abs(float(g["ll"])) + 1e-4
    for gk, pk in KEYMAP.items():
        assert rel(r["grads"][pk], g[gk]) < 2e-4, (gk, rel(r["grads"][pk], g[gk]))


@pytest.mark.parametrize("case", CASES + ["supervised_decode"])
def test_viterbi(golden, case):
    g = golden(case)
    mo = from_golden(g)
    spans, best, aux = mo.viterbi(g["features"], g["lengths"], g.get("valid_classes"), golden_addl_ends(g), g.get("constraints"))
    ref = g["viterbi_spans"]
    lenp = O.clamp_len_table(aux["lenp"], g["features"].shape[1])
    for b, T in enumerate(g["lengths"]):
        if (spans[b] == ref[b]).all():
            continue
        # allowed to differ only on a numerical near-tie of the fp32 reference
        n_cls = aux["em"].shape[-1]
        table = {int(c): i for i, c in enumerate(g["valid_classes"])} if "valid_classes" in g else {i: i for i in range(n_cls)}
        loc = np.array([table.get(int(x), -1) for x in ref[b, :T]])
        segs = O.segments_from_spans(loc, T)
        end = None if aux["ends"] is None else aux["ends"][b]
        s_ref = O.path_score(segs, aux["em"][b, :T], aux["init"], aux["trans"], lenp, end)
        assert best[b] - s_ref < 1e-4 * max(1.0, abs(best[b])), (case, b)


def test_known_answer(golden):
    """models/test_semimarkov.py:266-323 through the factorised oracle."""
    g = golden("known_answer")
    b, C, N, K, step = (int(g[k]) for k in ("b", "C", "N", "K", "step"))
    lengths = g["lengths_unpadded"]
    padded = N + 2 * step
    em = np.full((b, padded, C), O.BIG_NEG)
    for n in range(padded):
        em[:, n, (n // step) % C] = 1
    init = np.full(C, O.BIG_NEG)
    init[0] = 0
    lenp = np.full((K, C), O.BIG_NEG)
    lenp[step] = 0
    spans, _ = O.batch_viterbi(em, lengths, init, np.zeros((C, C)), lenp)
    assert (spans == g["sequence"]).all()


def test_labels_spans(golden):
    g = golden("labels_spans")
    assert (O.labels_to_spans(g["labels"], 10) == g["spans_k10"]).all()
    assert (O.spans_to_labels(g["spans_k10"]) == g["labels"]).all()
    for k in (2, 3, 5, 50):
        sp = O.labels_to_spans(g["rand_labels"], k)
        assert (sp == g["rand_spans_k%d" % k]).all()
        assert (O.spans_to_labels(sp) == g["rand_labels"]).all()
    assert (O.labels_to_spans(np.zeros((1, 6), dtype=np.int64), 4) == g["zeros_k4"]).all()
    rle = [[(0, 1), (1, 2), (2, 3)], [(0, 1), (1, 1), (2, 1), (3, 2), (4, 1)]]
    assert O.rle_spans(g["spans_k10"], [6, 6]) == rle
    assert O.rle_spans(g["spans_k10"], [5, 6])[0] == [(0, 1), (1, 2), (2, 2)]


def test_brute_force_agrees():
    rng = np.random.default_rng(3)
    for trial in range(4):
        T, C, K = int(rng.integers(2, 7)), int(rng.integers(1, 4)), int(rng.integers(2, 5))
        em = rng.normal(size=(T, C)) * 2
        init, trans = rng.normal(size=C), rng.normal(size=(C, C))
        lenp = rng.normal(size=(K, C))
        end = np.where(rng.random(C) < 0.3, O.BIG_NEG, 0.0)
        if T > (K - 1) * 10:
            continue
        bf = O.brute_force(em, init, trans, lenp, end)
        r = O.expected_counts(em, init, trans, lenp, end)
        v, segs = O.viterbi(em, init, trans, lenp, end)
        assert abs(bf["logZ"] - r["logZ"]) < 1e-9 * max(1, abs(bf["logZ"]))
        assert abs(bf["best"] - v) < 1e-9 * max(1, abs(v))
        if bf["gap"] > 1e-9:
            assert segs == bf["best_segs"]
        if bf["logZ"] > -1e8:
            for k in ("E_init", "E_trans", "E_len", "E_em"):
                assert np.allclose(bf[k], r[k], atol=1e-9), k


def test_materialised_equals_factorised():
    rng = np.random.default_rng(5)
    B, Tm, C, K = 3, 9, 3, 5
    lengths = np.array([9, 5, 7])
    em = rng.normal(size=(B, Tm, C))
    init, trans, lenp = rng.normal(size=C), rng.normal(size=(C, C)), rng.normal(size=(K, C))
    sc = O.log_hsmm(trans, em, init, lenp, lengths)
    z_m = O.materialised_dp(sc, lengths + 1, "log")
    v_m = O.materialised_dp(sc, lengths + 1, "max")
    z_f, _ = O.batch_logz_and_counts(em, lengths, init, trans, lenp)
    _, v_f = O.batch_viterbi(em, lengths, init, trans, lenp)
    assert np.allclose(z_m, z_f, rtol=1e-12) and np.allclose(v_m, v_f, rtol=1e-12)


def test_reference_port_matches_oracle():
    """oracle/reference_port.py (the bench's stand-in when oracle/_ref is absent): its logZ mean, parameter gradients and
    Viterbi spans against the factorised fp64 oracle on a small chain-constrained ragged batch."""
    import torch
    from oracle.module_oracle import ModuleOracle
    from oracle.reference_port import ReferencePort
    torch.manual_seed(0)
    C, D, K, B, T = 5, 6, 7, 3, 22
    means = torch.randn(C, D) * 0.5
    cov = torch.rand(D) + 0.5
    tl, il, lr = torch.randn(C, C) * 0.3, torch.rand(C), torch.log(torch.rand(C) * 3 + 1)
    tmask = torch.ones(C, C, dtype=torch.bool)
    for c in range(C):
        tmask[c, c] = False
        if c + 1 < C:
            tmask[c + 1, c] = False
    imask = torch.ones(C, dtype=torch.bool)
    imask[0] = False
    lengths = torch.LongTensor([22, 15, 9])
    lab = torch.sort(torch.randint(0, C, (B, T)), dim=1)[0]
    X = means[lab] + torch.randn(B, T, D)
    for b in range(B):
        X[b, int(lengths[b]):] = 0
    rp = ReferencePort(means, cov, tl, il, lr, K, tmask, imask)
    ends = [[C - 1] for _ in range(B)]
    ll, grads = rp.train_step(X, lengths, allowed_ends=ends)
    spans = rp.viterbi(X, lengths, allowed_ends=ends)
    params = dict(gaussian_means=means.numpy(), gaussian_cov=torch.diag(cov).numpy(), transition_logits=tl.numpy(),
                  init_logits=il.numpy(), poisson_log_rates=lr.numpy())
    mo = ModuleOracle(params, K, init_constraints=imask.numpy(), transition_constraints=tmask.numpy(), allowed_ends={C - 1})
    r = mo.log_likelihood(X.numpy(), lengths.numpy())
    assert abs(ll - r["ll"]) <= 1e-4 * abs(r["ll"])
    for k, g in grads.items():
        assert np.abs(g.numpy() - r["grads"][k]).max() <= 2e-3 * max(1e-9, np.abs(r["grads"][k]).max()), k
    o_spans, _, _ = mo.viterbi(X.numpy(), lengths.numpy())
    for b in range(B):
        n = int(lengths[b])
        assert (spans[b, :n + 1].numpy() == o_spans[b, :n + 1]).all()
