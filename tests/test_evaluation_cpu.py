"""CPU: the package's vectorised segmentation metrics (evaluation.py) against the reference's own `Accuracy` class
(evaluation/accuracy.py, imported unmodified), and the constraint construction of `SemiMarkovModel.from_args`,
`make_additional_allowed_ends` and `expand_constraints` against the reference wrapper (models/semimarkov/semimarkov.py:33-157)."""
import copy

import numpy as np
import pytest
import torch

import action_segmentation_b200 as pkg
from action_segmentation_b200 import data, evaluation
from action_segmentation_b200.args import HsmmArgs
from tests.golden import ref_import

needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="reference sources not present")


def _random_case(rng, n_videos, multi):
    G, P = [], []
    for v in range(n_videos):
        gt = np.repeat([0, 1, 0, 2, 0, 3, 4], rng.integers(3, 30, size=7))
        k = int(rng.integers(4, 9))
        pr = np.repeat(rng.integers(0, 5, size=k), rng.integers(3, 40, size=k))
        T = min(len(gt), len(pr))
        gt, pr = gt[:T], pr[:T]
        G.append([[int(x)] if not (multi and x == 1 and i % 3 == 0) else [1, 3] for i, x in enumerate(gt)])
        P.append(pr)
    return G, P


@needs_ref
@pytest.mark.parametrize("bkg,multi,optimal", [([0], False, False), ([0, 2], True, False), ([0], True, True), ([0], False, True)])
def test_metrics_identical_to_reference_accuracy(bkg, multi, optimal):
    Accuracy = ref_import.load_reference_module('evaluation.accuracy').Accuracy

    class Corpus:
        _background_indices = bkg
        index2label = {i: str(i) for i in range(16)}

    rng = np.random.default_rng(7)
    for trial in range(3):
        G, P = _random_case(rng, 5, multi)
        acc = Accuracy(verbose=False, corpus=Corpus())
        for g, p in zip(G, P):
            acc.add_gt_labels(g)
            acc.add_predicted_labels(list(p))
        acc.mof(optimal, possible_gt_labels=list(range(5)))
        acc.mof_classes()
        acc.iou_classes()
        acc.levenshtein()
        np.random.seed(trial)
        acc.single_step_recall()
        ref = acc.stat()
        np.random.seed(trial)
        mine = evaluation.segmentation_metrics(G, P, bkg, optimal_assignment=optimal)
        main_keys = ['mof', 'mof_non_bg', 'step_recall_non_bg', 'mean_normed_levenshtein', 'center_step_recall_non_bg', 'f1',
                     'f1_non_bg', 'pred_background', 'iou_multi_non_bg', 'predicted_label_types_per_video',
                     'predicted_label_types_non_bg_per_video', 'predicted_segments_per_video',
                     'predicted_segments_non_bg_per_video', 'multiple_gt_labels']  # main.py:20-26 STAT_KEYS
        for k in main_keys + ['mof_bg', 'precision', 'recall', 'single_step_recall', 'center_step_recall', 'total_levenshtein']:
            assert np.allclose(np.asarray(mine[k], dtype=float), np.asarray(ref[k], dtype=float), equal_nan=True), (k, mine[k], ref[k])


def test_edit_distance_known_answers():
    assert evaluation.edit_distance([1, 2, 3], [1, 2, 3]) == 0
    assert evaluation.edit_distance([], [1, 2]) == 2
    assert evaluation.edit_distance([1, 2, 3, 4], [2, 3]) == 2
    assert evaluation.edit_distance([0, 1, 0, 2, 0], [0, 2, 0, 1, 0]) == 2
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = list(rng.integers(0, 4, size=rng.integers(0, 12))), list(rng.integers(0, 4, size=rng.integers(0, 12)))
        assert evaluation.edit_distance(a, b) == ref_import._levenshtein(a, b)


@needs_ref
def test_from_args_constraints_match_reference_wrapper():
    ref_sm = ref_import.load_reference_module('models.semimarkov.semimarkov')
    split = data.make_crosstask_like(n_tasks=3, steps_per_task=(2, 5), n_videos=9, feature_dim=6, frames=(3, 30), narration=True, seed=4,
                                     allow_short=True)
    args = HsmmArgs(sm_max_span_length=8, sm_constrain_transitions=True, annotate_background_with_previous=True,
                    sm_constrain_with_narration=['train', 'test'], cuda=False, training='unsupervised')
    torch.manual_seed(0)
    mine = pkg.SemiMarkovModel.from_args(args, split)
    torch.manual_seed(0)
    ref = ref_sm.SemiMarkovModel.from_args(copy.copy(args), split)
    assert mine.model.merge_classes == ref.model.merge_classes
    assert mine.ordered_indices_by_task == ref.ordered_indices_by_task
    assert mine.model.allowed_ends == ref.model.allowed_ends
    for k, v in ref.model.state_dict().items():
        assert torch.equal(mine.model.state_dict()[k], v), k
    loader = data.make_data_loader(args, split, shuffle=False, batch_by_task=True, batch_size=3)
    short = 0
    for batch in loader:
        tasks, lengths = batch['task_name'], batch['lengths']
        a = mine.make_additional_allowed_ends(tasks, lengths)
        b = ref.make_additional_allowed_ends(tasks, lengths)
        assert a == b
        short += sum(1 for x in a if x)
        ea = mine.expand_constraints(split, tasks[0], batch['task_indices'][0], 1 - batch['constraints'])
        eb = ref.expand_constraints(split, tasks[0], batch['task_indices'][0], 1 - batch['constraints'])
        assert torch.equal(ea, eb)
    assert short > 0, "the split should contain a video shorter than its task chain"
    # no merging with --no_merge_classes, no constraints without the flag
    args2 = HsmmArgs(sm_max_span_length=8, annotate_background_with_previous=True, no_merge_classes=True, cuda=False)
    m2 = pkg.SemiMarkovModel.from_args(args2, split)
    assert m2.model.merge_classes is None and m2.model.transition_constraints is None and m2.ordered_indices_by_task is None
