"""CPU-side checks: the C-ABI library loads and exports every symbol include/hsmm_b200.h declares,
host-side span utilities match the reference's golden vectors, the data stand-in honours the batch
contract, and the product path refuses to run without CUDA (no fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

import action_segmentation_b200 as pkg
from action_segmentation_b200.args import HsmmArgs as RefArgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    header = open(os.path.join(ROOT, "include", "hsmm_b200.h")).read()
    declared = set(re.findall(r"\b(hsmm_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    lib = pkg._lib.load()
    for name in declared:
        assert hasattr(lib, name), "libhsmm_b200.so does not export %s" % name
    assert set(pkg._lib.EXPORTS) == declared
    assert lib.hsmm_version() >= 100


def test_library_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", pkg._lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_variant_table_covers_baseline_shapes():
    for C, K in [(23, 20), (9, 20), (11, 100), (23, 100), (48, 200), (48, 500)] + \
            [(c, k) for c in (16, 64, 133) for k in (50, 100, 200)]:
        assert pkg._lib.dp_variant(C, K, 0) != "unsupported", (C, K)
    for C, K in [(23, 20), (9, 20), (23, 100), (11, 100)]:
        for sparse in (False, True):
            assert pkg._lib.dp_variant(C, K, 1, sparse) != "unsupported"
            assert pkg._lib.dp_variant(C, K, 2, sparse) != "unsupported"
    assert "trans-sparse" in pkg._lib.dp_variant(23, 20, 1, True)


def test_sparse_transition_lists():
    allowed = torch.zeros(5, 5, dtype=torch.bool)
    for c in range(5):
        allowed[c, c] = True
        if c + 1 < 5:
            allowed[c + 1, c] = True
    pred, succ = pkg.hsmm.sparse_transition_lists(allowed, torch.device("cpu"))
    assert pred.tolist() == [[0, -1, -1, -1], [0, 1, -1, -1], [1, 2, -1, -1], [2, 3, -1, -1], [3, 4, -1, -1]]
    assert succ.tolist() == [[0, 1, -1, -1], [1, 2, -1, -1], [2, 3, -1, -1], [3, 4, -1, -1], [4, -1, -1, -1]]
    assert pkg.hsmm.sparse_transition_lists(torch.ones(6, 6, dtype=torch.bool), torch.device("cpu")) is None


def test_span_utils_golden(golden):
    g = golden("labels_spans")
    u = pkg.semimarkov_utils
    labels = torch.from_numpy(g["labels"])
    assert (u.labels_to_spans(labels, 10).numpy() == g["spans_k10"]).all()
    assert (u.spans_to_labels(torch.from_numpy(g["spans_k10"])).numpy() == g["labels"]).all()
    rand = torch.from_numpy(g["rand_labels"])
    for k in (2, 3, 5, 50):
        sp = u.labels_to_spans(rand, k)
        assert (sp.numpy() == g["rand_spans_k%d" % k]).all()
        assert (u.spans_to_labels(sp) == rand).all()
    assert (u.labels_to_spans(torch.zeros(1, 6).long(), 4).numpy() == g["zeros_k4"]).all()
    sp = torch.from_numpy(g["spans_k10"])
    assert u.rle_spans(sp, torch.LongTensor([6, 6])) == [[(0, 1), (1, 2), (2, 3)], [(0, 1), (1, 1), (2, 1), (3, 2), (4, 1)]]
    assert u.rle_spans(sp, torch.LongTensor([5, 6]))[0] == [(0, 1), (1, 2), (2, 2)]


def test_span_count_stats_matches_loop_definition():
    """Counting half of semimarkov_sufficient_stats against the oracle's span utilities."""
    from oracle import hsmm_oracle as O
    rng = np.random.default_rng(0)
    C, K = 4, 5
    labels = [torch.from_numpy(np.repeat(rng.integers(0, C, size=8), rng.integers(1, 9, size=8))) for _ in range(6)]
    st = pkg.semimarkov_utils.span_count_stats(labels, C, K)
    cnt = np.zeros(C); ln = np.zeros(C); start = np.zeros(C); tr = np.zeros((C, C))
    for lab in labels:
        sp = O.labels_to_spans(lab.numpy()[None], K)
        rle = O.rle_spans(sp, [lab.numel()])[0]
        start[rle[0][0]] += 1
        for i, (s, n) in enumerate(rle):
            cnt[s] += 1; ln[s] += n
            if i:
                tr[s, rle[i - 1][0]] += 1
    assert (st["span_counts"] == cnt).all() and (st["span_lengths"] == ln).all()
    assert (st["span_start_counts"] == start).all() and (st["span_transition_counts"] == tr).all()
    assert st["instance_count"] == 6


def test_parameter_transforms_match_oracle():
    """log_softmax / Poisson transforms of the module (CPU tensors are fine for these tiny ops)."""
    from oracle import hsmm_oracle as O
    m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=7), 5, 3, allow_self_transitions=True,
                             allowed_starts={0}, allowed_transitions={0: {0, 1}, 1: {1, 2}, 2: {2, 3}, 3: {3, 4}, 4: {4}},
                             allowed_ends={4}, merge_classes={0: 0, 1: 1, 2: 0, 3: 3, 4: 0})
    with torch.no_grad():
        m.transition_logits.normal_(); m.poisson_log_rates.normal_()
    vc = torch.LongTensor([0, 1, 2, 4])
    assert np.allclose(m.initial_log_probs(vc).detach().numpy(),
                       O.initial_log_probs(m.init_logits.detach().numpy(), m.init_constraints.numpy(), vc.numpy()), atol=1e-3, rtol=1e-6)
    assert np.allclose(m.transition_log_probs(vc).detach().numpy(),
                       O.transition_log_probs(m.transition_logits.detach().numpy(), m.transition_constraints.numpy(), vc.numpy()),
                       atol=1e-3, rtol=1e-6)
    merged = [0, 1, 0, 0]
    assert np.allclose(m.length_log_probs(vc).detach().numpy(),
                       O.poisson_length_log_probs(m.poisson_log_rates.detach().numpy()[merged], 7), atol=1e-5)
    assert set(m.state_dict().keys()) == {"poisson_log_rates", "gaussian_means", "gaussian_cov", "transition_logits",
                                          "init_logits", "init_constraints", "transition_constraints"}


def test_no_cpu_fallback():
    m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=5), 3, 4, allow_self_transitions=True)
    feats, lengths = torch.randn(2, 9, 4), torch.LongTensor([9, 7])
    with pytest.raises(pkg.HsmmError):
        m.log_likelihood(feats, lengths, None)
    with pytest.raises(pkg.HsmmError):
        m.viterbi(feats, lengths, None)
    with pytest.raises(pkg.HsmmError):
        pkg.hsmm.emission_scores(feats, torch.zeros(3, 4), torch.ones(4), None, lengths.int())


def test_data_stand_in_batch_contract():
    from action_segmentation_b200 import data
    split = data.make_crosstask_like(n_tasks=2, n_videos=6, feature_dim=8, narration=True, seed=1)
    starts, trans, ends, order = split.get_allowed_starts_and_transitions()
    assert len(starts) == 2 and len(ends) == 2
    loader = data.make_data_loader(RefArgs(), split, shuffle=True, batch_by_task=True, batch_size=2)
    seen = 0
    for batch in loader:
        assert len(set(batch["task_name"])) == 1
        B, T, D = batch["features"].shape
        assert batch["gt_single"].shape == (B, T) and batch["constraints"].shape[:2] == (B, T)
        assert int(batch["lengths"].max()) == T
        for i in range(B):
            assert (batch["features"][i, int(batch["lengths"][i]):] == 0).all()
            assert batch["task_indices"][i].tolist() == order[batch["task_name"][i]]
        seen += B
    assert seen == 6


def test_em_update_closed_form_on_given_statistics():
    """M-step arithmetic (no kernels involved): normalised counts under the constraint masks, mean segment length,
    class means = wx / wsum, untouched parameters for classes without mass."""
    from action_segmentation_b200.args import HsmmArgs as RefArgs
    C, D = 4, 3
    m = pkg.SemiMarkovModule(RefArgs(sm_max_span_length=10), C, D, allow_self_transitions=True, allowed_starts={0},
                             allowed_transitions={0: {0, 1}, 1: {1, 2}, 2: {2, 3}, 3: {3}}, allowed_ends={3})
    old_means = m.gaussian_means.detach().clone()
    stats = dict(wx=torch.tensor([[2.0, 4.0, 6.0], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [3.0, 0.0, 3.0]]),
                 wsum=torch.tensor([2.0, 4.0, 0.0, 1.0]), init=torch.tensor([5.0, 0.0, 0.0, 0.0]),
                 trans=torch.tensor([[3.0, 9.0, 0, 0], [1.0, 2.0, 0, 0], [0, 2.0, 0.0, 0], [0, 0, 1.0, 4.0]]),
                 len_num=torch.tensor([12.0, 6.0, 0.0, 5.0]), len_den=torch.tensor([4.0, 3.0, 0.0, 1.0]),
                 logz=torch.tensor([-50.0]), n=torch.tensor([5.0]))
    ll = m.em_update(stats)
    assert ll == -10.0
    assert torch.allclose(m.gaussian_means[0], torch.tensor([1.0, 2.0, 3.0]))
    assert torch.allclose(m.gaussian_means[2], old_means[2])  # no mass: untouched
    assert torch.allclose(m.poisson_log_rates[:2].exp(), torch.tensor([3.0, 2.0]))
    p = m.transition_log_probs(None).exp()  # [to, from], masked + normalised as the DP sees it
    assert torch.allclose(p[:, 0], torch.tensor([0.75, 0.25, 0.0, 0.0]), atol=1e-6)  # the masked count 9.0 at [0,1] is ignored
    assert torch.allclose(p[:, 1], torch.tensor([0.0, 0.5, 0.5, 0.0]), atol=1e-6)
    assert torch.allclose(m.initial_log_probs(None).exp(), torch.tensor([1.0, 0.0, 0.0, 0.0]), atol=1e-6)
    buf = m.pack_statistics(stats)
    assert buf.numel() == C * D + C + C + C * C + C + C + 2
    assert all(torch.equal(m.unpack_statistics(buf)[k], stats[k]) for k in m.STAT_KEYS)
