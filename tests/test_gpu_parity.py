"""GPU parity tests: the CUDA path (through the C-ABI) against the oracle and the golden vectors."""
import numpy as np
import pytest
import torch

from oracle import hsmm_oracle as O
from oracle.module_oracle import from_golden, golden_addl_ends
from tests.helpers import (check_viterbi_against_oracle, module_from_golden, random_problem, rel_err, sparse_lists,
                           to_dev)

pytestmark = pytest.mark.gpu

CASES = ["unconstrained", "short_clamp", "constrained", "constrained_narration"]
KEYMAP = dict(g_means="gaussian_means", g_trans="transition_logits", g_init="init_logits", g_rates="poisson_log_rates")


def _inputs(g):
    feats = torch.from_numpy(g["features"]).cuda()
    lengths = torch.from_numpy(g["lengths"]).long()
    B = feats.shape[0]
    vpi = None
    if "valid_classes" in g:
        vpi = [torch.from_numpy(g["valid_classes"]).long() for _ in range(B)]
    cons = torch.from_numpy(g["constraints"]).cuda() if "constraints" in g else None
    return feats, lengths, vpi, golden_addl_ends(g), cons


@pytest.mark.parametrize("case", CASES)
def test_emission_golden(golden, case):
    g = golden(case)
    m = module_from_golden(g)
    feats, lengths, vpi, addl, cons = _inputs(g)
    elp = m.emission_log_probs(feats, None if vpi is None else vpi[0], cons).cpu().numpy()
    for b, T in enumerate(g["lengths"]):
        assert rel_err(elp[b, :T], g["elp"][b, :T]) < 1e-5


@pytest.mark.parametrize("B,Tmax,D,C,pen", [(7, 300, 200, 23, False), (5, 257, 200, 13, True), (3, 700, 64, 48, False),
                                            (4, 130, 200, 7, True), (2, 1000, 300, 64, False), (3, 90, 36, 5, False),
                                            (300, 40, 200, 23, False), (97, 300, 64, 9, True),
                                            # C > 64: one launch per block of 64 classes + the finishing kernel
                                            (5, 300, 200, 133, False), (4, 257, 64, 65, True), (3, 200, 200, 284, False),
                                            (60, 70, 128, 130, True)])
def test_emission_tensor_core_vs_oracle(B, Tmax, D, C, pen):
    """hsmm_emission on the tcgen05/TMA path (3xTF32) against the fp64 oracle (semimarkov_modules.py:324-381)
    and against the SIMT fp32 kernel: same em/rowterm/offset contract, fp32-level accuracy."""
    import action_segmentation_b200 as pkg
    assert pkg._lib.load().hsmm_emission_workspace_bytes(D, C) > 0
    rng = np.random.default_rng(B * 1000 + D + C)
    lengths = rng.integers(1, Tmax + 1, size=B)
    lengths[0] = Tmax
    if B > 50:
        lengths[3::11] = 0   # many videos (several windows of the live-tile cursor), some of them empty
    means = rng.normal(size=(C, D)) * 0.5
    cov = rng.uniform(0.5, 1.5, size=D)
    lab = rng.integers(0, C, size=(B, Tmax))
    X = (means[lab] + rng.normal(size=(B, Tmax, D))).astype(np.float32)
    for b in range(B):
        X[b, lengths[b]:] = 0
    penalty = None
    if pen:
        penalty = np.where(rng.uniform(size=(B, Tmax, C)) < 0.2, -1e4, 0.0).astype(np.float32)
        penalty[:, :, 0] = 0
    dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    f32 = lambda a: torch.from_numpy(a.astype(np.float32)).cuda()  # noqa: E731
    li = torch.from_numpy(lengths).to(torch.int32).cuda()
    outs = []
    for tc in (True, False):
        em, rowterm, offset = pkg.hsmm.emission_scores(dev(X), f32(means), f32(cov), dev(penalty), li, tensor_cores=tc)
        outs.append((em.cpu().numpy().astype(np.float64), rowterm.cpu().numpy().astype(np.float64), offset.cpu().numpy()))
    m32 = means.astype(np.float32).astype(np.float64)
    c32 = cov.astype(np.float32).astype(np.float64)
    ref = O.emission_log_probs(X.astype(np.float64), m32, c32, None if penalty is None else penalty.astype(np.float64))
    for em, rowterm, offset in outs:
        assert em.shape[2] == pkg.hsmm.ldc_of(C)
        elp = em[:, :, :C] + rowterm[:, :, None]
        for b, T in enumerate(lengths):
            if T == 0:
                assert (em[b] == 0).all() and (rowterm[b] == 0).all() and offset[b] == 0
                continue
            scale = np.abs(ref[b, :T]).max()
            assert np.abs(elp[b, :T] - ref[b, :T]).max() < 2e-6 * scale + 1e-3 * (1 if pen else 0), (b, np.abs(elp[b, :T] - ref[b, :T]).max())
            assert em[b, :T, :C].max(axis=1).max() <= 1e-6 and (em[b, :T, :C].max(axis=1) > -1e-3).all()  # best class 0
            assert (em[b, T:] == 0).all() and (rowterm[b, T:] == 0).all()
            assert abs(offset[b] - rowterm[b, :T].sum()) < 1e-6 * abs(offset[b]) + 1e-6
    # the two kernels agree on the class-relative scores to fp32 rounding of the dot products
    d = np.abs(outs[0][0] - outs[1][0])
    unpen = np.ones_like(d, dtype=bool) if penalty is None else np.pad(penalty == 0, ((0, 0), (0, 0), (0, d.shape[2] - C)))
    assert d[unpen].max() < 4e-6 * np.abs(ref).max(), d[unpen].max()


@pytest.mark.parametrize("case", CASES)
def test_loglik_and_grads_golden(golden, case, pair_mode):
    """logZ, its batch mean and the four parameter gradients: within 1e-4 relative of the reference's
    fp32 values, and tighter against the fp64 oracle."""
    g = golden(case)
    m = module_from_golden(g)
    feats, lengths, vpi, addl, cons = _inputs(g)
    ll, log_det = m.log_likelihood(feats, lengths, vpi, spans=None, add_eos=True,
                                   additional_allowed_ends_per_instance=addl, constraints=cons)
    ll.backward()
    assert abs(float(ll) - float(g["ll"])) <= 1e-4 * abs(float(g["ll"]))
    assert float(log_det) == 0.0
    r = from_golden(g).log_likelihood(g["features"], g["lengths"], g.get("valid_classes"), addl, g.get("constraints"))
    assert abs(float(ll) - r["ll"]) <= 2e-6 * abs(r["ll"])
    for gk, pk in KEYMAP.items():
        mine = getattr(m, pk).grad.cpu().numpy()
        # bar: 1e-4 relative against the fp64 oracle.  Where fp32 itself cannot resolve the inputs (the
        # -1e4 narration penalty rounds a penalised emission to ~1e-3 absolute) the reference's own fp32
        # result misses the oracle by more than that; then we must be at least as close as 3x its error.
        ref_noise = rel_err(g[gk], r["grads"][pk])
        assert rel_err(mine, r["grads"][pk]) < max(1e-4, 3 * ref_noise), (gk, rel_err(mine, r["grads"][pk]), ref_noise)
        assert rel_err(mine, g[gk]) < 1e-4 + 3 * ref_noise, (gk, rel_err(mine, g[gk]), ref_noise)


@pytest.mark.parametrize("case", CASES + ["supervised_decode"])
def test_viterbi_golden(golden, case, pair_mode):
    g = golden(case)
    m = module_from_golden(g)
    feats, lengths, vpi, addl, cons = _inputs(g)
    spans, labels = m.viterbi(feats, lengths, vpi, add_eos=True, additional_allowed_ends_per_instance=addl,
                              constraints=cons, return_labels=True)
    assert spans.dtype == torch.int64 and not spans.is_cuda
    ref = g["viterbi_spans"]
    if not (spans.numpy() == ref).all():
        # only a numerical near-tie may differ: compare fp64 path scores through the oracle
        o_spans, best, aux = from_golden(g).viterbi(g["features"], g["lengths"], g.get("valid_classes"), addl, g.get("constraints"))
        prob = dict(em=aux["em"], lengths=g["lengths"], init=aux["init"], trans=aux["trans"], lenp=aux["lenp"], end=aux["ends"])
        n_cls = m.n_classes
        table = {int(c): i for i, c in enumerate(g["valid_classes"])} if "valid_classes" in g else {i: i for i in range(n_cls)}
        table[n_cls] = aux["em"].shape[-1]
        table[-1] = -1
        local = np.vectorize(table.get)(spans.numpy())
        check_viterbi_against_oracle(prob, local, em_round=False)
    # per-frame labels agree with the span encoding
    import action_segmentation_b200 as pkg
    lab_ref = pkg.semimarkov_utils.spans_to_labels(spans)
    for b, T in enumerate(g["lengths"]):
        assert (labels[b, :T] == lab_ref[b, :T]).all()


def test_known_answer(golden):
    """Body of the reference's test_log_hsmm (models/test_semimarkov.py:266-323) through hsmm_viterbi."""
    import action_segmentation_b200 as pkg
    g = golden("known_answer")
    b, C, N, K, step = (int(g[k]) for k in ("b", "C", "N", "K", "step"))
    padded = N + 2 * step
    em = np.full((b, padded, C), O.BIG_NEG)
    for n in range(padded):
        em[:, n, (n // step) % C] = 1
    init = np.full(C, O.BIG_NEG)
    init[0] = 0
    lenp = np.full((K, C), O.BIG_NEG)
    lenp[step] = 0
    prob = dict(em=em, lengths=g["lengths_unpadded"], init=init, trans=np.zeros((C, C)), lenp=lenp, end=None)
    d = to_dev(prob)
    spans, labels, score = pkg.hsmm.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], None, None,
                                                   d["lengths_i32"], d["order"])
    assert (spans.cpu().numpy() == g["sequence"]).all()
    for s in range(N // step):
        assert (spans[:, step * s].cpu() == s % C).all()


SHAPES = [
    # (B, Tmax, C, K, chain, ends)  -- one per compiled DP variant / layout regime
    (9, 60, 23, 20, True, True),     # reg<20,1> trans-reg  (flagship CrossTask shape)
    (9, 60, 9, 20, True, True),      # reg<10,2> trans-reg
    (7, 150, 11, 100, False, False),  # reg<50,2> trans-reg, one warp (the S6 shape)
    (5, 150, 23, 100, True, True),   # reg<25,4> trans-smem, 3 warps
    (5, 120, 16, 50, False, False),  # reg<25,2> trans-reg
    (4, 120, 64, 50, False, False),  # reg<25,2> trans-smem 4 warps
    (3, 90, 133, 50, False, False),  # reg<25,2> 9 warps
    (3, 220, 133, 200, False, False),  # reg<50,4> len-smem 17 warps
    (3, 260, 64, 200, False, False),  # reg<25,8> 16 warps
    (2, 520, 48, 500, False, False),  # reg<63,8> len-smem
    (6, 40, 3, 30, False, False),    # reg<32,1>
    (6, 70, 5, 52, False, False),    # reg<13,4>
    (4, 30, 1, 5, False, False),     # single class
    (5, 12, 4, 40, False, False),    # K clamped to the padded length
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "B%d_T%d_C%d_K%d" % s[:4])
def test_viterbi_random_vs_oracle(shape, pair_mode):
    import action_segmentation_b200 as pkg
    B, Tmax, C, K, chain, ends = shape
    rng = np.random.default_rng(100 + C * 7 + K)
    prob = random_problem(rng, B, Tmax, C, K, Tmin=1, chain=chain, ends=ends)
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    sp = sparse_lists(prob) if chain else None  # ordering-constrained shapes run the sparse-transition kernels
    spans, labels, score = pkg.hsmm.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None,
                                                   d["lengths_i32"], d["order"], trans_pred=None if sp is None else sp[0])
    n_exact = check_viterbi_against_oracle(prob, spans.cpu().numpy(), score.cpu().numpy())
    assert n_exact >= B - 1, "more than one video differs from the oracle path (%d of %d exact)" % (n_exact, B)
    if chain:  # the hint must not change anything: dense kernels on the same inputs
        spans_d, _, score_d = pkg.hsmm.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None,
                                                      d["lengths_i32"], d["order"])
        assert (spans_d == spans).all()
        assert torch.allclose(score_d, score, rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("f64_state", [False, True], ids=["f32state", "f64state"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "B%d_T%d_C%d_K%d" % s[:4])
def test_logz_and_counts_random_vs_oracle(shape, f64_state, pair_mode):
    """logZ within 1e-5 relative, expected counts within 1e-4 relative of the fp64 oracle.  With the f64-state
    kernels (HSMM_FLAG_F64_STATE, what the module selects whenever narration constraints are given) the
    ordering-constrained shapes also carry -1e4 narration penalties on (nearly) every path -- inputs on which
    the reference's own fp32 algorithm is only good to ~1e-3 (measured with oracle/reference_port.py:
    E_trans 3e-3, E_em 7e-3 on the C=23 case)."""
    import action_segmentation_b200 as pkg
    B, Tmax, C, K, chain, ends = shape
    rng = np.random.default_rng(200 + C * 7 + K)
    prob = random_problem(rng, B, Tmax, C, K, Tmin=1, chain=chain, ends=ends, narration=chain and f64_state)
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    sp = sparse_lists(prob) if chain else (None, None)
    logz, saved = pkg.hsmm.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"],
                                        d["order"], trans_pred=sp[0], f64_state=f64_state)
    w = rng.uniform(0.5, 1.5, size=B)
    g = torch.from_numpy(w).float().cuda()
    d_init, d_trans, d_len, d_em = pkg.hsmm.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"],
                                                         d["lengths_i32"], d["order"], g, saved, trans_succ=sp[1],
                                                         f64_state=f64_state)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    p32 = dict(prob, em=f32(prob["em"]), init=f32(prob["init"]), trans=f32(prob["trans"]), lenp=f32(prob["lenp"]))
    ref_logz, acc = O.batch_logz_and_counts(p32["em"], prob["lengths"], p32["init"], p32["trans"], p32["lenp"], prob["end"], w)
    assert np.allclose(logz.cpu().numpy(), ref_logz, rtol=1e-5, atol=1e-4)
    mine = dict(E_init=d_init, E_trans=d_trans, E_len=d_len, E_em=d_em[:, :, :C])
    for k, v in mine.items():
        assert rel_err(v.cpu().numpy(), acc[k]) < 1e-4, (k, rel_err(v.cpu().numpy(), acc[k]))


@pytest.mark.parametrize("C,K", [(9, 20), (23, 20), (23, 100)])
def test_sparse_hint_degenerate_falls_back_to_dense(C, K, pair_mode):
    """Videos with NO path through the unmasked transitions (chain longer than the video, no extra
    allowed end): the sparse kernels must notice (result <= -1e8) and reproduce the dense answer, which
    itself must agree with the oracle on the -1e9-penalised problem."""
    import action_segmentation_b200 as pkg
    rng = np.random.default_rng(7 + C + K)
    B, Tmax = 6, 40
    prob = random_problem(rng, B, Tmax, C, K, Tmin=20, chain=True, ends=False)
    prob["lengths"][1] = 3   # cannot reach the last class of the chain: every path pays -1e9
    prob["lengths"][4] = 5
    end = np.full((B, C), O.BIG_NEG)
    end[:, C - 1] = 0.0
    prob["end"] = end
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    pred, succ = sparse_lists(prob)
    g = torch.ones(B, device="cuda")
    outs = []
    for hint in ((pred, succ), (None, None)):
        logz, saved = pkg.hsmm.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"],
                                            d["order"], trans_pred=hint[0])
        grads = pkg.hsmm.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"],
                                       g, saved, trans_succ=hint[1])
        spans, _, score = pkg.hsmm.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None,
                                                  d["lengths_i32"], d["order"], trans_pred=hint[0])
        outs.append((logz, grads, spans, score))
    (lz_s, gr_s, sp_s, sc_s), (lz_d, gr_d, sp_d, sc_d) = outs
    assert float(lz_d[1]) < -1e8 and float(lz_d[4]) < -1e8 and float(lz_d[0]) > -1e8
    assert torch.allclose(lz_s, lz_d, rtol=1e-6)
    assert torch.allclose(sc_s, sc_d, rtol=1e-6)
    good = [0, 2, 3, 5]
    for b in good:
        assert (sp_s[b] == sp_d[b]).all()
    # feasible videos: identical frame posteriors whatever the hint
    assert torch.allclose(gr_s[3][good], gr_d[3][good], rtol=1e-4, atol=1e-5)
    # the two degenerate videos carry -1e9 on every path (fp32 resolves ~1e2 there): their posteriors and
    # hence the summed counts agree to fp32-at-1e9 noise only
    for a, b_ in zip(gr_s, gr_d):
        assert torch.allclose(a, b_, rtol=2e-3, atol=2e-3)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    ref_logz, _ = O.batch_logz_and_counts(f32(prob["em"]), prob["lengths"], f32(prob["init"]), f32(prob["trans"]),
                                          f32(prob["lenp"]), prob["end"])
    assert np.allclose(lz_d.cpu().numpy(), ref_logz, rtol=1e-6)


@pytest.mark.parametrize("B,Tmax,D,C", [(6, 333, 200, 23), (5, 200, 200, 7), (3, 150, 64, 48), (2, 90, 300, 13), (4, 77, 6, 5),
                                        (2, 130, 200, 133)])
def test_weighted_feature_sums_vs_numpy(B, Tmax, D, C):
    """hsmm_weighted_feature_sums (the reduction behind d/d gaussian_means and the supervised class means,
    semimarkov_utils.py:74-126) against a float64 numpy contraction."""
    import action_segmentation_b200 as pkg
    rng = np.random.default_rng(B + Tmax + D + C)
    lengths = rng.integers(1, Tmax + 1, size=B)
    lengths[0] = Tmax
    X = rng.normal(size=(B, Tmax, D)).astype(np.float32)
    ldc = pkg.hsmm.ldc_of(C)
    wgt = np.zeros((B, Tmax, ldc), dtype=np.float32)
    wgt[:, :, :C] = rng.dirichlet(np.ones(C), size=(B, Tmax))
    li = torch.from_numpy(lengths).to(torch.int32).cuda()
    wx, wsum = pkg.hsmm.weighted_feature_sums(torch.from_numpy(X).cuda(), torch.from_numpy(wgt).cuda(), C, li)
    ref_wx = np.zeros((C, D))
    ref_ws = np.zeros(C)
    for b, T in enumerate(lengths):
        ref_wx += wgt[b, :T, :C].astype(np.float64).T @ X[b, :T].astype(np.float64)
        ref_ws += wgt[b, :T, :C].astype(np.float64).sum(axis=0)
    scale = np.sqrt(float(lengths.sum()))  # typical magnitude of a sum of ~N(0, w^2) terms
    assert np.abs(wx.cpu().numpy() - ref_wx).max() < 2e-5 * scale
    assert np.abs(wsum.cpu().numpy() - ref_ws).max() < 1e-5 * np.abs(ref_ws).max()


@pytest.mark.parametrize("B,Tmax,D,C", [(3, 100, 224, 32), (4, 130, 128, 16), (2, 64, 4, 1), (5, 257, 200, 23), (3, 90, 228, 9),
                                        (2, 70, 200, 33), (300, 45, 200, 23), (130, 70, 204, 40), (40, 150, 300, 23), (9, 120, 300, 48),
                                        (4, 200, 896, 96), (3, 64, 900, 12), (3, 64, 64, 97), (7, 100, 452, 65)])
def test_weighted_feature_sums_tensor_core_edges(B, Tmax, D, C):
    """The tcgen05 path of hsmm_weighted_feature_sums at the edges of its eligibility (one launch per block of 224 features x 32
    classes, up to D = 896 and C = 96: the README's D = 300 features and Breakfast's 48 classes run here; D = 900 and C = 97
    fall to the SIMT kernel), with a zero-length video, and with NaN in every padding frame of the features AND of
    the weights: frames t >= length must not reach the sums (the reference only ever sums the first length_b frames,
    semimarkov_utils.py:74-126)."""
    import action_segmentation_b200 as pkg
    rng = np.random.default_rng(7 * B + Tmax + D + C)
    lengths = rng.integers(1, Tmax + 1, size=B)
    lengths[0] = Tmax
    if B > 2:
        lengths[1] = 0
    if B > 50:
        lengths[5::13] = 0   # several windows of the live-tile cursor, with empty videos
    X = rng.normal(size=(B, Tmax, D)).astype(np.float32)
    ldc = pkg.hsmm.ldc_of(C)
    wgt = np.zeros((B, Tmax, ldc), dtype=np.float32)
    wgt[:, :, :C] = rng.dirichlet(np.ones(C), size=(B, Tmax))
    ref_wx = np.zeros((C, D))
    ref_ws = np.zeros(C)
    for b, T in enumerate(lengths):
        ref_wx += wgt[b, :T, :C].astype(np.float64).T @ X[b, :T].astype(np.float64)
        ref_ws += wgt[b, :T, :C].astype(np.float64).sum(axis=0)
        X[b, T:] = np.nan
        wgt[b, T:] = np.nan
    li = torch.from_numpy(lengths).to(torch.int32).cuda()
    wx, wsum = pkg.hsmm.weighted_feature_sums(torch.from_numpy(X).cuda(), torch.from_numpy(wgt).cuda(), C, li)
    scale = np.sqrt(float(lengths.sum()))
    assert np.isfinite(wx.cpu().numpy()).all() and np.isfinite(wsum.cpu().numpy()).all()
    assert np.abs(wx.cpu().numpy() - ref_wx).max() < 2e-5 * scale
    assert np.abs(wsum.cpu().numpy() - ref_ws).max() < 1e-5 * np.abs(ref_ws).max()


def test_supervised_fit_golden(golden):
    """fit_supervised closed form (semimarkov_modules.py:195-256) against the reference's fitted parameters."""
    import action_segmentation_b200 as pkg
    from action_segmentation_b200.args import HsmmArgs as RefArgs
    g = golden("supervised_fit")
    C, K = int(g["n_classes"]), int(g["max_k"])
    D = g["features"].shape[1]
    m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=K), C, D, allow_self_transitions=True).cuda()
    feats, labels, off = [], [], 0
    for n in g["lengths"]:
        feats.append(torch.from_numpy(g["features"][off:off + n]))
        labels.append(torch.from_numpy(g["labels"][off:off + n]))
        off += n
    m.fit_supervised(feats, labels)
    for k in ("gaussian_means", "gaussian_cov", "transition_logits", "init_logits", "poisson_log_rates"):
        assert rel_err(getattr(m, k).detach().cpu().numpy(), g[k]) < 2e-5, k


def test_gold_score_golden(golden):
    """log_likelihood with gold spans, generative and discriminative (semimarkov_modules.py:626-655)."""
    g = golden("gold_score")
    for tag, disc in (("gen", False), ("disc", True)):
        m = module_from_golden(g, sm_train_discriminatively=disc)
        feats = torch.from_numpy(g["features"]).cuda()
        lengths = torch.from_numpy(g["lengths"]).long()
        spans = torch.from_numpy(g["spans"]).long()
        ll, _ = m.log_likelihood(feats, lengths, None, spans=spans, add_eos=True)
        ll.backward()
        assert abs(float(ll) - float(g["ll_" + tag])) <= 1e-4 * abs(float(g["ll_" + tag])), tag
        for gk, pk in KEYMAP.items():
            mine = getattr(m, pk).grad.cpu().numpy()
            assert rel_err(mine, g[gk + "_" + tag]) < 1e-4, (tag, gk)


def test_full_size_properties():
    """BASELINE-size batch (CrossTask shape: D=200, C=23, K=20, T up to 3000): size-independent
    invariants of the DP outputs."""
    import action_segmentation_b200 as pkg
    from action_segmentation_b200.args import HsmmArgs as RefArgs
    torch.manual_seed(5)
    B, Tmax, D, C, K = 64, 3000, 200, 23, 20
    m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=K), C, D, allow_self_transitions=True).cuda()
    with torch.no_grad():
        m.gaussian_means.normal_(0, 0.3)
        m.poisson_log_rates.uniform_(1.0, 2.5)
        m.transition_logits.normal_()
    lengths = torch.randint(1000, Tmax + 1, (B,))
    lengths[0] = Tmax
    lab = torch.randint(0, C, (B, Tmax // 10 + 1)).repeat_interleave(10, dim=1)[:, :Tmax].cuda()
    feats = m.gaussian_means.detach()[lab] + torch.randn(B, Tmax, D, device="cuda")
    s = m._scores(feats, lengths, None, None)
    em, rowterm, offset = pkg.hsmm.emission_scores(feats, s["means"], s["cov_diag"], None, s["lengths_i32"])
    assert float(em[:, :, :C].max()) <= 0.0 + 1e-6  # shifted by the per-frame best class
    logz, saved = pkg.hsmm.logz_forward(em, C, s["init"], s["trans"], s["lenp"], None, offset, s["lengths_i32"], s["order"])
    g = torch.ones(B, device="cuda")
    d_init, d_trans, d_len, d_em = pkg.hsmm.logz_backward(em, C, s["init"].detach(), s["trans"].detach(), s["lenp"].detach(),
                                                         None, s["lengths_i32"], s["order"], g, saved)
    spans, labels, score = pkg.hsmm.viterbi_decode(em, C, s["init"], s["trans"], s["lenp"], None, offset,
                                                   s["lengths_i32"], s["order"])
    lens_dev = lengths.cuda()
    mask = (torch.arange(Tmax, device="cuda")[None] < lens_dev[:, None]).float()
    # posterior over classes sums to one on every real frame, zero on padding
    assert torch.allclose(d_em[:, :, :C].sum(-1), mask, atol=2e-4)
    # one first segment per video; expected segment lengths add up to the number of frames
    assert abs(float(d_init.sum()) - B) < 1e-3 * B
    k = torch.arange(K, device="cuda", dtype=torch.float32)[:, None]
    assert abs(float((d_len * k).sum()) - float(lengths.sum())) < 1e-4 * float(lengths.sum())
    # segments = transitions + one per video
    assert abs(float(d_len.sum()) - float(d_trans.sum()) - B) < 1e-4 * float(d_len.sum())
    # max-plus <= log-sum; decoded labels agree with spans; spans only start valid segments
    assert (score <= logz + 1e-6 * logz.abs()).all()
    lab_ref = pkg.semimarkov_utils.spans_to_labels(spans)
    assert (torch.where(mask.bool(), labels, 0) == torch.where(mask.bool(), lab_ref[:, :Tmax], 0)).all()
    assert (spans[torch.arange(B), lens_dev] == C).all()
    # fp64 score of the decoded path equals the kernel's score (three videos, oracle as checker)
    em_h, off_h = em.cpu().numpy().astype(np.float64), offset.cpu().numpy()
    for b in (0, 1, B - 1):
        T = int(lengths[b])
        segs = O.segments_from_spans(spans[b].cpu().numpy(), T)
        sc = O.path_score(segs, em_h[b, :T, :C], s["init"].detach().cpu().numpy().astype(np.float64),
                          s["trans"].detach().cpu().numpy().astype(np.float64), s["lenp"].detach().cpu().numpy().astype(np.float64))
        assert abs(sc + off_h[b] - float(score[b])) < 1e-6 * abs(float(score[b]))


# ---------------------------------------------------------------------------------------------
# linear-window (block floating point) kernels vs the log-domain kernels
# ---------------------------------------------------------------------------------------------
def _saved_flags(saved, B, Tmax, C):
    """(fflag, bflag) of the `saved` buffer of hsmm_logz_forward (layout: hsmm_api.cu `carve`)."""
    import action_segmentation_b200 as pkg
    plane = B * (Tmax + 1) * pkg.hsmm.ldc_of(C)
    off = B * 8 + 2 * plane * 4 + B * (Tmax + 1) * 4
    f = saved[off:off + 8 * B].view(torch.float32).cpu().numpy()
    return f[:B], f[B:]


def _fwd_bwd(prob, d, sp, w):
    import action_segmentation_b200 as pkg
    C = prob["em"].shape[2]
    logz, saved = pkg.hsmm.logz_forward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"],
                                        d["order"], trans_pred=sp[0])
    fflag = _saved_flags(saved, prob["em"].shape[0], prob["em"].shape[1], C)[0].copy()
    g = torch.from_numpy(w).float().cuda()
    grads = pkg.hsmm.logz_backward(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], d["lengths_i32"], d["order"], g,
                                   saved, trans_succ=sp[1])
    torch.cuda.synchronize()
    bflag = _saved_flags(saved, prob["em"].shape[0], prob["em"].shape[1], C)[1].copy()
    return logz, grads, fflag, bflag


LIN_SHAPES = [
    # (B, Tmax, C, K, chain, ends, em scale)
    (9, 300, 23, 20, True, True, 3.0),    # lin<20,1> sparse (flagship)
    (9, 300, 13, 20, True, True, 12.0),   # lin<10,2> sparse, peaked emissions
    (7, 200, 23, 20, False, False, 3.0),  # lin<20,1> dense transitions in registers
    (7, 200, 16, 20, False, False, 25.0),  # lin<10,2> dense
    (6, 150, 5, 52, False, False, 3.0),   # lin<13,4>
    (6, 150, 16, 50, False, False, 3.0),  # lin<25,2>
    (6, 120, 3, 30, False, False, 3.0),   # lin<32,1>
    (6, 150, 11, 100, False, False, 3.0),  # lin<50,2> (the S6 shape: C <= 16, windows of 51..100 frames)
    (5, 140, 16, 77, True, True, 6.0),    # lin<50,2> sparse
]


@pytest.mark.parametrize("shape", LIN_SHAPES, ids=lambda s: "B%d_T%d_C%d_K%d_s%g" % (s[0], s[1], s[2], s[3], s[6]))
def test_linear_window_matches_log_domain_and_oracle(shape, pair_mode):
    """Same inputs through the linear-window kernels and (hsmm_set_linear_window(0)) the log-domain kernels: logZ,
    every expected count and the frame posteriors agree to fp32 rounding, and both agree with the fp64 oracle."""
    import action_segmentation_b200 as pkg
    B, Tmax, C, K, chain, ends, scale = shape
    rng = np.random.default_rng(300 + C * 7 + K)
    prob = random_problem(rng, B, Tmax, C, K, Tmin=Tmax // 3, chain=chain, ends=ends, scale=scale)
    # realistic duration rates (the reference initialises Poisson rate 1 and fits mean segment lengths)
    prob["lenp"] = O.poisson_length_log_probs(np.log(rng.uniform(1.0, 12.0, size=C)), K)
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    sp = sparse_lists(prob) if chain else (None, None)
    w = rng.uniform(0.5, 1.5, size=B)
    assert pkg._lib.dp_variant(C, K, 1, chain).startswith("lin+")
    try:
        lz1, g1, ff1, bf1 = _fwd_bwd(prob, d, sp, w)
        pkg._lib.set_linear_window(False)
        assert not pkg._lib.dp_variant(C, K, 1, chain).startswith("lin+")
        lz0, g0, ff0, bf0 = _fwd_bwd(prob, d, sp, w)
    finally:
        pkg._lib.set_linear_window(True)
    assert (ff0 < 2).all()  # (bflag is only written by the linear-window backward kernel)
    if K <= 32:
        # ordinary inputs with short windows are certified by the linear-window kernels: nothing goes to the fallback
        assert (ff1 < 2).all() and (bf1 == 0).all(), (ff1, bf1)
    else:
        # Poisson tails below 2^-100 inside a long window: the length-table check (reason bit 8) hands every video over
        assert (ff1 >= 4).all() and (bf1 == 9).all(), (ff1, bf1)
    assert torch.allclose(lz1, lz0, rtol=2e-6, atol=5e-4)
    for a, b_ in zip(g1, g0):  # two fp32 computations, each within 1e-4 of the oracle (checked below)
        assert rel_err(a.cpu().numpy(), b_.cpu().numpy()) < 2.5e-4
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    ref_logz, acc = O.batch_logz_and_counts(f32(prob["em"]), prob["lengths"], f32(prob["init"]), f32(prob["trans"]),
                                            f32(prob["lenp"]), prob["end"], w)
    assert np.allclose(lz1.cpu().numpy(), ref_logz, rtol=1e-5, atol=1e-4)
    mine = dict(E_init=g1[0], E_trans=g1[1], E_len=g1[2], E_em=g1[3][:, :, :C])
    for k, v in mine.items():
        assert rel_err(v.cpu().numpy(), acc[k]) < 1e-4, (k, rel_err(v.cpu().numpy(), acc[k]))


@pytest.mark.parametrize("case", ["tiny_rates", "no_self_loops", "no_path", "collapse"])
def test_linear_window_flags_fall_back_to_log_domain(case, pair_mode):
    """Inputs the float window cannot certify: the videos are flagged, recomputed by the log-domain kernels, and the
    results still match the oracle."""
    import action_segmentation_b200 as pkg
    rng = np.random.default_rng(17)
    B, Tmax, C, K = 6, 60, 9, 20
    prob = random_problem(rng, B, Tmax, C, K, Tmin=30, chain=True, ends=True, scale=3.0)
    expect_all = False
    if case == "tiny_rates":      # usable lengths with probability below 2^-100: the window could overflow
        prob["lenp"] = O.poisson_length_log_probs(np.log(np.full(C, 0.02)), K)
        expect_all = True
    elif case == "no_self_loops":  # strict chain c -> c+1 only and peaked emissions: a class sum collapses when its
        logits = np.full((C, C), O.BIG_NEG)  # only mass leaves the window
        for c in range(C - 1):
            logits[c + 1, c] = 0.0
        logits[C - 1, C - 1] = 0.0
        prob["trans"] = O.log_softmax(logits, axis=0)
        prob["em"] = prob["em"] * 10.0
        prob["lengths"][:] = rng.integers(3 * C, Tmax + 1, size=B)
    elif case == "no_path":      # chain longer than the video: every path pays -1e9
        prob["lengths"][1] = 4
        prob["lengths"][3] = 6
        end = np.full((B, C), O.BIG_NEG)
        end[:, C - 1] = 0.0
        prob["end"] = end
    elif case == "collapse":     # entry mass into a class decays by 2^-40 per frame, then everything older expires
        prob = random_problem(rng, B, Tmax, 5, K, Tmin=50, chain=False, ends=False, scale=1.0)
        C = 5
        prob["em"][:, :, 1] = -np.arange(Tmax)[None, :] * 28.0 * (np.arange(Tmax)[None, :] < 12)
        prob["trans"][1, :] = -40.0   # hard to enter class 1 ...
        prob["init"] = np.log(np.array([0.01, 0.96, 0.01, 0.01, 0.01]))  # ... except at the start
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    d = to_dev(prob)
    chain = case != "collapse"
    sp = sparse_lists(prob) if chain else (None, None)
    w = np.ones(B)
    assert pkg._lib.dp_variant(C, K, 1, chain).startswith("lin+")
    lz, g, ff, bf = _fwd_bwd(prob, d, sp, w)
    n_flagged = int((ff >= 4).sum())
    if expect_all:
        assert n_flagged == B and (bf >= 1).all()
    if case == "no_path":
        assert ff[1] >= 4 and ff[3] >= 4
    f32 = lambda x: x.astype(np.float32).astype(np.float64)  # noqa: E731
    ref_logz, acc = O.batch_logz_and_counts(f32(prob["em"]), prob["lengths"], f32(prob["init"]), f32(prob["trans"]),
                                            f32(prob["lenp"]), prob["end"], w)
    assert np.allclose(lz.cpu().numpy(), ref_logz, rtol=1e-5, atol=1e-4), (case, n_flagged)
    if case != "no_path":  # (the -1e9-penalised videos resolve their counts to fp32-at-1e9 noise only)
        mine = dict(E_init=g[0], E_trans=g[1], E_len=g[2], E_em=g[3][:, :, :C])
        for k, v in mine.items():
            assert rel_err(v.cpu().numpy(), acc[k]) < 1e-4, (case, k, rel_err(v.cpu().numpy(), acc[k]), n_flagged)
    print(case, "flagged forward:", n_flagged, "backward:", int((bf > 0).sum()))


VIT2_SHAPES = [
    # (B, Tmax, C, K, chain, ends, em scale)
    (16, 400, 23, 20, True, True, 3.0),   # vit2<20,1> sparse (flagship)
    (16, 400, 13, 20, True, True, 12.0),  # vit2<10,2> sparse
    (12, 300, 23, 20, False, False, 3.0),  # vit2<20,1> dense
    (12, 300, 11, 20, False, False, 1.0),  # vit2<10,2> dense, many near ties
    (8, 200, 5, 30, False, False, 3.0),   # vit2<13,4>
    (8, 200, 16, 33, False, False, 3.0),  # vit2<25,2>, L = 32
    (8, 150, 3, 30, False, False, 3.0),   # vit2<32,1>
    (6, 12, 4, 40, False, False, 3.0),    # K clamped to the padded length
]


@pytest.mark.parametrize("shape", VIT2_SHAPES, ids=lambda s: "B%d_T%d_C%d_K%d_s%g" % (s[0], s[1], s[2], s[3], s[6]))
def test_viterbi_deferred_argmax_identical_to_generic_kernel(shape, pair_mode):
    """The deferred-arg-max kernel rebuilds the sweep's candidates bit for bit during the back-trace: spans, labels
    and scores must be IDENTICAL to dp_forward_kernel<VIT> (hsmm_set_linear_window(0)), and match the oracle."""
    import action_segmentation_b200 as pkg
    B, Tmax, C, K, chain, ends, scale = shape
    rng = np.random.default_rng(500 + C * 7 + K)
    prob = random_problem(rng, B, Tmax, C, K, Tmin=1, chain=chain, ends=ends, scale=scale)
    if chain:  # one video too short for the chain: no path through the sparse hint -> flagged -> generic kernel
        prob["lengths"][2] = 3
        end = np.full((B, C), O.BIG_NEG)
        end[:, C - 1] = 0.0
        prob["end"] = end
    prob["lenp"] = O.clamp_len_table(prob["lenp"], Tmax)
    K_eff = prob["lenp"].shape[0]
    d = to_dev(prob)
    sp = sparse_lists(prob) if chain else None
    cid = torch.arange(100, 100 + C + 1, dtype=torch.int32, device="cuda")  # global class ids

    def run():
        return pkg.hsmm.viterbi_decode(d["em"], C, d["init"], d["trans"], d["lenp"], d["end"], None, d["lengths_i32"],
                                       d["order"], class_ids=cid, trans_pred=None if sp is None else sp[0])
    assert pkg._lib.dp_variant(C, K_eff, 0, chain).startswith("lin+")
    try:
        s1, l1, sc1 = run()
        pkg._lib.set_linear_window(False)
        s0, l0, sc0 = run()
    finally:
        pkg._lib.set_linear_window(True)
    assert (s1 == s0).all() and (l1 == l0).all()
    assert torch.allclose(sc1, sc0, rtol=1e-6, atol=1e-6)  # (the normaliser sum is grouped differently)
    # against the oracle (local ids)
    spans_local = torch.where(s1 >= 100, s1 - 100, s1).cpu().numpy()
    feas = [b for b in range(B) if not (chain and prob["lengths"][b] < C)]
    sub = dict(prob, em=prob["em"][feas], lengths=prob["lengths"][feas], end=None if prob["end"] is None else prob["end"][feas])
    # long runs of one class can be cut into the same multiset of lengths in several orders with EXACTLY the same
    # score (self-transitions), so the oracle's tie-break need not be ours: require optimality to fp32 rounding
    # (fp64 score of the decoded path within 1e-6 relative of the oracle's best), not the identical path
    check_viterbi_against_oracle(sub, spans_local[feas], sc1.cpu().numpy()[feas], tol=1e-6)


def test_em_statistics_and_closed_form_update():
    """`expected_statistics` (the E-step behind SURVEY section 8f item 3) against the oracle's expected counts in global
    class layout, and the closed-form M-step: every EM iteration must not decrease the mean log-likelihood."""
    import action_segmentation_b200 as pkg
    from oracle.module_oracle import ModuleOracle
    from action_segmentation_b200.args import HsmmArgs as RefArgs
    torch.manual_seed(11)
    B, T, D, C, K = 12, 120, 16, 7, 20
    chain = {c: {c, c + 1} for c in range(C - 1)}
    chain[C - 1] = {C - 1}
    m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=K), C, D, allow_self_transitions=True, allowed_starts={0},
                             allowed_transitions=chain, allowed_ends={C - 1}).cuda()
    true_means = torch.randn(C, D, device="cuda") * 1.5
    with torch.no_grad():
        m.gaussian_means.copy_(true_means + 0.8 * torch.randn(C, D, device="cuda"))
        m.poisson_log_rates.fill_(np.log(6.0))
        m.transition_logits.normal_(0, 0.3)
    lengths = torch.randint(60, T + 1, (B,))
    lengths[0] = T
    lab = torch.stack([torch.sort(torch.randint(0, C, (T,)))[0] for _ in range(B)]).cuda()
    feats = true_means[lab] + torch.randn(B, T, D, device="cuda")
    for i, n in enumerate(lengths):
        feats[i, n:] = 0

    stats = m.expected_statistics(feats, lengths, None)
    params = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
    mo = ModuleOracle(params, K, init_constraints=params["init_constraints"],
                      transition_constraints=params["transition_constraints"], allowed_ends={C - 1})
    r = mo.log_likelihood(feats.cpu().numpy(), lengths.numpy())
    assert abs(float(stats["logz"]) / B - r["ll"]) <= 1e-5 * abs(r["ll"])
    # per-video normalisation and frame bookkeeping
    assert abs(float(stats["init"].sum()) - B) < 1e-3 * B
    assert abs(float(stats["wsum"].sum()) - float(lengths.sum())) < 1e-4 * float(lengths.sum())
    assert abs(float(stats["len_num"].sum()) - float(lengths.sum())) < 1e-4 * float(lengths.sum())
    assert abs(float(stats["len_den"].sum()) - float(stats["trans"].sum()) - B) < 1e-3 * float(stats["len_den"].sum())
    # d logZ / d means = (wx - wsum mu) / var must reproduce the oracle's gradient (mean over the batch)
    var = torch.diagonal(m.gaussian_cov)
    g_means = (stats["wx"] - stats["wsum"][:, None] * m.gaussian_means.detach()) / var[None, :] / B
    assert rel_err(g_means.cpu().numpy(), r["grads"]["gaussian_means"]) < 1e-4
    # pack / unpack (what one all-reduce carries)
    back = m.unpack_statistics(m.pack_statistics(stats))
    assert all(torch.equal(back[k], stats[k]) for k in m.STAT_KEYS)

    lls = []
    for it in range(4):
        lls.append(m.em_update(m.expected_statistics(feats, lengths, None)))
    final = float(m.expected_statistics(feats, lengths, None)["logz"]) / B
    lls.append(final)
    for a, b_ in zip(lls, lls[1:]):
        assert b_ >= a - 1e-4 * abs(a), lls
    assert lls[-1] > lls[0] + 1.0, lls
    # the means moved towards the generating ones
    assert float((m.gaussian_means.detach() - true_means).norm()) < 0.5 * float((0.8 * torch.ones(C, D)).norm())


@pytest.mark.parametrize("B,Tmax,W", [(5, 300, 200), (3, 77, 24), (40, 130, 64), (2, 9000, 200)])
@pytest.mark.parametrize("mapped", [False, True])
def test_upload_ragged_copies_exactly_the_live_rows(B, Tmax, W, mapped):
    """hsmm_upload_ragged / hsmm_upload_ragged_mapped: the device buffer receives the rows t < lengths[b] of the padded
    host batch (models/model.py:42-63 `padding_colate` + `.cuda()`), bit for bit, and keeps what it held behind them."""
    import action_segmentation_b200 as pkg
    rng = np.random.default_rng(B + Tmax + W)
    lengths = rng.integers(0, Tmax + 1, size=B)
    lengths[0] = Tmax
    if B > 2:
        lengths[1] = 0
    host = torch.from_numpy(rng.normal(size=(B, Tmax, W)).astype(np.float32)).pin_memory()
    dev = torch.full((B, Tmax, W), -7.0, device="cuda")
    lh = torch.from_numpy(lengths).to(torch.int32)
    n = pkg.hsmm.upload_ragged(host, dev, lh, lh.cuda() if mapped else None)
    torch.cuda.synchronize()
    assert n == int(lengths.sum()) * W * 4
    got = dev.cpu()
    for b, T in enumerate(lengths):
        assert torch.equal(got[b, :T], host[b, :T])
        assert (got[b, T:] == -7.0).all()
    if mapped:  # pageable host memory is refused by the kernel path, not silently copied
        lib = pkg._lib.load()
        pageable = torch.zeros(B, Tmax, W)
        rc = lib.hsmm_upload_ragged_mapped(pageable.data_ptr(), dev.data_ptr(), lh.cuda().data_ptr(), B, Tmax, W, None)
        assert rc != 0
