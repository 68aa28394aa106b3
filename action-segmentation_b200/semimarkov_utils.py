"""Span/label conversions and supervised sufficient statistics -- same names and results as the
reference's models/semimarkov/semimarkov_utils.py, without its Python loops over frames.

    labels_to_spans   semimarkov_utils.py:6-23
    rle_spans         semimarkov_utils.py:26-48
    spans_to_labels   semimarkov_utils.py:51-63
    semimarkov_sufficient_stats  semimarkov_utils.py:74-126  (feature reductions run on the GPU
                      through hsmm_weighted_feature_sums / hsmm_feature_moments)
"""
import numpy as np
import torch


def labels_to_spans(position_labels, max_k):
    """b x N labels -> span encoding (class id at segment starts, -1 inside); runs are split every
    max_k - 1 frames so that every span has a usable length."""
    assert not (position_labels == -1).any(), "position_labels already appear span encoded (have -1)"
    b, N = position_labels.shape
    pos = torch.arange(N, device=position_labels.device).unsqueeze(0).expand(b, N)
    change = torch.ones_like(position_labels, dtype=torch.bool)
    change[:, 1:] = position_labels[:, 1:] != position_labels[:, :-1]
    run_start = torch.cummax(torch.where(change, pos, torch.zeros_like(pos)), dim=1)[0]
    r = pos - run_start
    if max_k is None:
        is_start = r == 0
    elif max_k - 1 <= 0:
        is_start = torch.ones_like(change)
    else:
        is_start = (r % (max_k - 1)) == 0
    return torch.where(is_start, position_labels, torch.full_like(position_labels, -1))


def spans_to_labels(spans):
    """b x N span encoding -> per-frame labels (continuations take the label of their span start)."""
    b, N = spans.shape
    assert (spans[:, 0] != -1).all()
    pos = torch.arange(N, device=spans.device).unsqueeze(0).expand(b, N)
    last_start = torch.cummax(torch.where(spans != -1, pos, torch.zeros_like(pos)), dim=1)[0]
    return torch.gather(spans, 1, last_start)


def rle_spans(spans, lengths):
    """[(symbol, count), ...] per row over the first lengths[i] positions."""
    spans = spans.detach().cpu().numpy()
    out = []
    for i in range(spans.shape[0]):
        row = spans[i, :int(lengths[i])]
        if row.size == 0:
            out.append([])
            continue
        starts = np.flatnonzero(row != -1)
        if starts.size == 0 or starts[0] != 0:
            starts = np.concatenate([[0], starts])
        counts = np.diff(np.concatenate([starts, [row.size]]))
        rle = [(int(row[s]), int(c)) for s, c in zip(starts, counts)]
        assert sum(c for _, c in rle) == int(lengths[i])
        out.append(rle)
    return out


def span_count_stats(label_list, n_classes, max_k):
    """Counting half of semimarkov_sufficient_stats (semimarkov_utils.py:84-111): span starts,
    span counts, summed span lengths and [to, from] transition counts (float32, as the reference)."""
    span_counts = np.zeros(n_classes, dtype=np.float32)
    span_lengths = np.zeros(n_classes, dtype=np.float32)
    span_start_counts = np.zeros(n_classes, dtype=np.float32)
    span_transition_counts = np.zeros((n_classes, n_classes), dtype=np.float32)
    for labels in label_list:
        lab = labels.detach().cpu().numpy() if isinstance(labels, torch.Tensor) else np.asarray(labels)
        if lab.size == 0:
            continue
        change = np.flatnonzero(np.concatenate([[True], lab[1:] != lab[:-1]]))
        run_len = np.diff(np.concatenate([change, [lab.size]]))
        syms, lens = [], []
        step = None if max_k is None else max(max_k - 1, 1)
        for s, ln in zip(lab[change], run_len):
            if step is None:
                syms.append(s)
                lens.append(ln)
            else:
                full, rem = divmod(int(ln), step)
                syms.extend([s] * (full + (1 if rem else 0)))
                lens.extend([step] * full + ([rem] if rem else []))
        syms = np.asarray(syms, dtype=np.int64)
        lens = np.asarray(lens, dtype=np.float32)
        span_start_counts[syms[0]] += 1
        np.add.at(span_counts, syms, 1)
        np.add.at(span_lengths, syms, lens)
        np.add.at(span_transition_counts, (syms[1:], syms[:-1]), 1)
    return {
        'span_counts': span_counts,
        'span_lengths': span_lengths,
        'span_start_counts': span_start_counts,
        'span_transition_counts': span_transition_counts,
        'instance_count': len(label_list),
    }
