"""Segmentation metrics of the reference's evaluation code, vectorised.

The reference scores predictions with `Accuracy` (/root/reference/src/evaluation/accuracy.py): `mof` (:475-579) walks
every frame of the split in a Python loop, `single_step_recall` (:410-472) and `levenshtein` (:364-408) loop over the
videos.  Once the DP runs on the GPU that loop is what an epoch waits for (the per-epoch
callback decodes and evaluates train + dev, main.py:207-218).  `segmentation_metrics` computes the same statistics --
same names, same [numerator, denominator] pairs `main.py:186-194` aggregates -- with numpy array operations.

Identity with the reference class (including the `np.random.choice` draw order of `step_recall_non_bg`, so even that
statistic agrees under the same numpy seed) is pinned by tests/test_evaluation_cpu.py against the unmodified class.
"""
from collections import defaultdict

import numpy as np


def run_length_encode(labels):
    """evaluation/accuracy.py:21-37 -> (symbols, counts) arrays."""
    labels = np.asarray(labels)
    if labels.size == 0:
        return labels, np.zeros(0, dtype=np.int64)
    starts = np.flatnonzero(np.concatenate([[True], labels[1:] != labels[:-1]]))
    return labels[starts], np.diff(np.concatenate([starts, [labels.size]]))


def edit_distance(a, b):
    """Levenshtein distance between two symbol sequences (the reference calls the `editdistance` C extension)."""
    a, b = list(a), np.asarray(list(b))
    prev = np.arange(len(b) + 1)
    for i, x in enumerate(a, 1):
        sub = prev[:-1] + (b != x)
        cur = np.minimum(prev[1:] + 1, sub)
        # insertions propagate left to right: cur[j] = min(cur[j], cur[j-1] + 1)
        cur = np.concatenate([[i], cur])
        cur = np.minimum.accumulate(cur - np.arange(len(b) + 1)) + np.arange(len(b) + 1)
        prev = cur
    return int(prev[-1])


def _assignment(gt, pred, optimal_assignment):
    """gt label -> [cluster] (accuracy.py:232-318, 334-345): identity, or Hungarian on the co-occurrence table."""
    gt2cluster = defaultdict(list)
    if not optimal_assignment:
        for label in np.unique(gt):
            gt2cluster[label] = [label]
        return gt2cluster
    from scipy.optimize import linear_sum_assignment
    ug, up = np.unique(gt), np.unique(pred)
    size = max(len(ug), len(up))
    gt_labels, pr_labels = list(ug), list(up)
    for lst in (gt_labels, pr_labels):  # pad with unused label ids, as _create_voting_table does
        idx = len(lst)
        while len(lst) < size:
            cand = idx
            while cand in lst:
                cand += 1
            lst.append(cand)
            idx += 1
    table = np.zeros((size, size))
    gi = np.searchsorted(ug, gt)
    pi = np.searchsorted(up, pred)
    np.add.at(table, (gi, pi), 1.0)
    x, y = linear_sum_assignment(-table)
    for a, b in zip(x, y):
        gt2cluster[gt_labels[a]] = [pr_labels[b]]
    return gt2cluster


def segmentation_metrics(gt_per_video, pred_per_video, background_indices, optimal_assignment=False):
    """gt_per_video: per video a list of per-frame label LISTS (`Video.gt()`: several labels may hold at a frame; the
    first is the single label) or a 1-D array of single labels; pred_per_video: per video a 1-D label array.
    Returns {statistic: np.array([numerator, denominator])} for every key of main.py's STAT_KEYS plus the other
    entries `Accuracy.stat()` carries ('mof_bg', 'precision', 'recall', ...)."""
    bkg = [int(x) for x in background_indices]
    single, multi_len, multi_sets = [], [], []
    for g in gt_per_video:
        if len(g) and isinstance(g[0], (list, tuple, np.ndarray)):
            single.append(np.asarray([t[0] for t in g]))
            multi_len.append(np.asarray([len(t) for t in g]))
            multi_sets.append(g)
        else:
            g = np.asarray(g)
            single.append(g)
            multi_len.append(np.ones(len(g), dtype=np.int64))
            multi_sets.append(None)
    preds = [np.asarray(p) for p in pred_per_video]
    assert len(single) == len(preds)
    for g, p in zip(single, preds):
        assert len(g) == len(p), "{} != {}".format(len(g), len(p))
    gt = np.concatenate(single)
    pred = np.concatenate(preds)
    nlab = np.concatenate(multi_len)
    gt2cluster = _assignment(gt, pred, optimal_assignment)

    def cluster_of(labels):
        """Remap gt labels to their cluster (-1: no cluster, i.e. an empty list in the reference)."""
        out = np.full(len(labels), -1, dtype=np.int64)
        for lab in np.unique(labels):
            cl = gt2cluster[lab] if lab in gt2cluster else []
            if len(cl) > 0:
                out[labels == lab] = cl[0]
        return out

    ret = {}
    # ---- mof / per-class counts (accuracy.py:475-520, 581-612) -----------------------------------------
    frames_true = 0.0
    tot_true = tot = tot_true_nb = tot_nb = 0.0
    for lab in np.unique(gt):
        mask = gt == lab
        true = 0.0
        for cl in gt2cluster[lab]:
            true += float(np.sum(pred[mask] == cl))
        frames_true += true
        tot_true += true
        tot += float(mask.sum())
        if lab not in bkg:
            tot_true_nb += true
            tot_nb += float(mask.sum())
    ret['mof'] = [frames_true, len(gt)]
    ret['mof_bg'] = [tot_true, tot]
    ret['mof_non_bg'] = [tot_true_nb, tot_nb]
    # ---- per-frame precision / recall with multiple gt labels (accuracy.py:522-577) -------------------
    # true positive: the prediction equals the cluster of ANY of the frame's gt labels
    tp = pred == cluster_of(gt)
    if (nlab > 1).any():
        off = 0
        for g, m in zip(single, multi_sets):
            if m is not None:
                for t in np.flatnonzero(np.asarray([len(x) for x in m]) > 1):
                    cl = [gt2cluster[x][0] for x in m[t] if x in gt2cluster and len(gt2cluster[x]) > 0]
                    tp[off + t] = pred[off + t] in cl
            off += len(g)
    bkg_clusters = [gt2cluster[b][0] for b in bkg if b in gt2cluster and len(gt2cluster[b]) > 0]
    pred_bkg = np.isin(pred, bkg_clusters)
    is_bkg = np.isin(gt, bkg)
    n = float(len(gt))
    precision = np.array([float(tp.sum()), n])
    recall = np.array([float(tp.sum()), float(nlab.sum())])
    ret['precision'], ret['recall'] = precision, recall
    p = precision[0] / precision[1] if precision[1] else 0.0
    r = recall[0] / recall[1] if recall[1] else 0.0
    ret['f1'] = np.array([(2 * p * r) / (p + r), 1.0]) if (p + r) > 0 else np.array([float('nan'), 1.0])
    nb = ~is_bkg
    p_nb = np.array([float(tp[nb].sum()), float(nb.sum())])
    r_nb = np.array([float(tp[nb].sum()), float(nlab[nb].sum())])
    ret['precision_non_bg'], ret['recall_non_bg'] = p_nb, r_nb
    pn = p_nb[0] / p_nb[1] if p_nb[1] else 0.0
    rn = r_nb[0] / r_nb[1] if r_nb[1] else 0.0
    ret['f1_non_bg'] = np.array([0 if (pn == 0 and rn == 0) else (2 * pn * rn) / (pn + rn), 1.0])
    ret['true_background'] = np.array([float(is_bkg.sum()), n])
    ret['pred_background'] = np.array([float(pred_bkg.sum()), n])
    either = (~is_bkg) | (~pred_bkg)
    ret['iou_multi_non_bg'] = np.array([float(tp[either].sum()), float(either.sum())])
    ret['multiple_gt_labels'] = np.array([float((nlab > 1).sum()), n])
    # ---- segment statistics (accuracy.py:364-408) ------------------------------------------------------
    bkg_remapped = set(bkg_clusters)
    levs, maxsegs = [], []
    pred_segments = pred_segments_nb = 0.0
    for g, pr in zip(single, preds):
        gs, _ = run_length_encode(g)
        ps, _ = run_length_encode(pr)
        gs_remapped = [gt2cluster[x][0] for x in gs]
        pred_segments += len(ps)
        pred_segments_nb += sum(1 for x in ps if x not in bkg_remapped)
        levs.append(edit_distance(gs_remapped, ps))
        maxsegs.append(max(len(gs_remapped), len(ps)))
    levs, maxsegs = np.array(levs, dtype=np.float64), np.array(maxsegs, dtype=np.float64)
    nv = float(len(preds))
    ret.update({
        'mean_levenshtein': np.array([np.mean(levs), 1.0]),
        'mean_max_segments': np.array([np.mean(maxsegs), 1.0]),
        'total_levenshtein': np.array([np.sum(levs), 1.0]),
        'num_videos': np.array([nv, 1.0]),
        'mean_normed_levenshtein': np.array([np.mean(levs / maxsegs), 1.0]),
        'predicted_segments_per_video': np.array([pred_segments, nv]),
        'predicted_segments_non_bg_per_video': np.array([pred_segments_nb, nv]),
    })
    # ---- step recall (accuracy.py:410-472); np.random.choice is drawn in the reference's order ---------------
    step_match = step_total = nb_match = nb_total = center_match = nb_center_match = 0.0
    types = types_nb = 0.0
    for g, pr in zip(single, preds):
        remapped = cluster_of(g)
        for lab in np.unique(pr):
            types += 1
            if lab not in bkg_remapped:
                types_nb += 1
        for lab in np.unique(remapped):
            step_total += 1
            non_bg = lab not in bkg_remapped
            if non_bg:
                nb_total += 1
            idx = np.flatnonzero(pr == lab)
            if len(idx) == 0:
                continue
            pick = np.random.choice(idx)
            mid = (idx[0] + idx[-1]) / 2
            center = idx[np.argmin(np.abs(idx - mid))]
            if remapped[pick] == lab:
                step_match += 1
                nb_match += non_bg
            if remapped[center] == lab:
                center_match += 1
                nb_center_match += non_bg
    ret.update({
        'single_step_recall': np.array([step_match, step_total]),
        'step_recall_non_bg': np.array([nb_match, nb_total]),
        'center_step_recall': np.array([center_match, step_total]),
        'center_step_recall_non_bg': np.array([nb_center_match, nb_total]),
        'predicted_label_types_per_video': np.array([types, nv]),
        'predicted_label_types_non_bg_per_video': np.array([types_nb, nv]),
    })
    return ret


def summarize(stats_by_task, keys):
    """main.py:186-194: sum the [numerator, denominator] pairs over tasks, then divide."""
    out = {}
    for key in keys:
        s = np.array([np.asarray(st[key], dtype=np.float64) for st in stats_by_task.values()]).sum(axis=0)
        out[key] = float(s[0]) / s[1]
    return out
