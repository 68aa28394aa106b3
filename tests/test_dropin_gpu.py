"""Drop-in proof (-m gpu): the B200 `SemiMarkovModel` driven the way the reference drives its own
(/root/reference/src/main.py:163-266: from_args -> fit with a callback that decodes and pickles every epoch ->
pickle.loads of the best model -> predict -> evaluation), next to the UNMODIFIED reference classes run on the CPU over
oracle/torch_struct_shim.py on the same data, same parameters.

  * S6 flow: closed-form supervised fit, parameters and predictions against the reference wrapper's;
  * U7 flow: --sm_constrain_transitions --annotate_background_with_previous --sm_constrain_with_narration train test,
    2 epochs of Adam on logZ; losses against the reference's fit, predictions against the reference's predict
    with the trained parameters loaded into its module;
  * frame accuracy (MoF) and step recall from the reference's `Accuracy` are IDENTICAL for both prediction sets;
  * the reference's own main.train / main.test functions drive the B200 classifier (the INTEGRATION.md swap).

The reference sources come from /root/reference or oracle/_ref (tests/golden/ref_import.py)."""
import argparse
import copy
import pickle

import numpy as np
import pytest
import torch

import action_segmentation_b200 as pkg
from action_segmentation_b200 import data
from action_segmentation_b200.args import HsmmArgs
from tests.golden import ref_import

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="reference sources not present (oracle/_ref)")


def reference_accuracy(split, predictions, optimal_assignment=False):
    """MoF / step recall of `predictions` from the reference's own Accuracy class, per task, the way
    Datasplit.accuracy_corpus feeds it (data/corpus.py:466-494, 547-576)."""
    acc_mod = ref_import.load_reference_module('evaluation.accuracy')
    by_task = {}
    for v in split.videos:
        by_task.setdefault(v['task_name'], []).append(v)
    out = {}
    for task, vids in by_task.items():
        # numpy >= 2.2 raises on `x in [[], ...]` for numpy scalars (accuracy.py:555 relied on the old "empty array is
        # False" rule): evaluate against the background indices that occur, which is what the old rule amounted to
        class _Corpus:
            pass
        corpus = _Corpus()
        canon = split.corpus.annotate_background_with_previous
        corpus._background_indices = [split.corpus._background_indices[0]] if canon else list(split.corpus._background_indices)
        corpus.index2label = split.corpus.index2label
        acc = acc_mod.Accuracy(verbose=False, corpus=corpus)
        for v in sorted(vids, key=lambda x: x['video_name']):
            gt = [int(x) for x in v['gt_single']]
            pred = [int(x) for x in predictions[v['video_name']]]
            if canon:
                gt = [split.canonicalize_background(x) for x in gt]
                pred = [split.canonicalize_background(x) for x in pred]
            acc.add_gt_labels([[x] for x in gt])
            acc.add_predicted_labels(pred)
        acc.mof(optimal_assignment, possible_gt_labels=split.corpus.indices_by_task(task))
        acc.mof_classes()
        acc.levenshtein()
        np.random.seed(0)
        acc.single_step_recall()
        out[task] = {k: np.asarray(v, dtype=np.float64) for k, v in acc.stat().items()}
    return out


def assert_same_metrics(split, pred_a, pred_b):
    a, b = reference_accuracy(split, pred_a), reference_accuracy(split, pred_b)
    for task in a:
        for key in ('mof', 'mof_non_bg', 'center_step_recall_non_bg', 'step_recall_non_bg', 'mean_normed_levenshtein',
                    'f1', 'pred_background'):
            assert np.array_equal(a[task][key], b[task][key]), (task, key, a[task][key], b[task][key])
    # and the package's vectorised metrics agree with the reference class on these predictions
    mine = split.accuracy_corpus(False, lambda video: pred_a[video.name], verbose=False)
    for task in a:
        for key in ('mof', 'mof_non_bg', 'center_step_recall_non_bg', 'mean_normed_levenshtein', 'f1'):
            assert np.allclose(np.asarray(mine[task][key], dtype=np.float64), a[task][key]), (task, key)
    return a


def frames_equal(pred_a, pred_b):
    same = total = 0
    for k in pred_a:
        same += int((np.asarray(pred_a[k]) == np.asarray(pred_b[k])).sum())
        total += len(pred_a[k])
    return same, total


@needs_ref
def test_s6_supervised_flow_matches_reference_wrapper():
    ref_sm = ref_import.load_reference_module('models.semimarkov.semimarkov')
    train = data.make_supervised_like(n_tasks=2, n_videos=10, feature_dim=12, frames=(60, 110), seed=3)
    test = data.make_supervised_like(n_tasks=2, n_videos=6, feature_dim=12, frames=(60, 110), seed=3)
    test.videos = [dict(v, video_name='t' + v['video_name']) for v in test.videos[::-1]]
    args = HsmmArgs(sm_max_span_length=20, training='supervised', batch_size=3)

    model = pkg.SemiMarkovModel.from_args(args, train)
    model.fit(train, use_labels=True)
    model = pickle.loads(pickle.dumps(model))  # main.py:234, 248-257
    pred = model.predict(test)

    rargs = copy.copy(args)
    rargs.cuda = False
    ref = ref_sm.SemiMarkovModel.from_args(rargs, train)
    ref.fit(train, use_labels=True)
    ref_pred = ref.predict(test)

    for k, v in ref.model.state_dict().items():
        mine = model.model.state_dict()[k].cpu()
        assert torch.allclose(mine, v, rtol=2e-5, atol=2e-6), (k, float((mine - v).abs().max()))
    assert set(pred) == set(ref_pred)
    same, total = frames_equal(pred, ref_pred)
    assert same == total, "%d of %d frames differ from the reference wrapper's predictions" % (total - same, total)
    stats = assert_same_metrics(test, pred, ref_pred)
    assert all(s['mof'][1] > 0 for s in stats.values())


@needs_ref
def test_u7_unsupervised_flow_matches_reference_wrapper():
    ref_sm = ref_import.load_reference_module('models.semimarkov.semimarkov')
    split = data.make_crosstask_like(n_tasks=3, steps_per_task=(2, 4), n_videos=12, feature_dim=10, frames=(40, 70),
                                     narration=True, seed=5)
    args = HsmmArgs(sm_max_span_length=10, sm_constrain_transitions=True, annotate_background_with_previous=True,
                    sm_constrain_with_narration=['train', 'test'], epochs=2, batch_size=3, training='unsupervised',
                    print_every=0)

    torch.manual_seed(11)
    model = pkg.SemiMarkovModel.from_args(args, split)
    init_state = {k: v.detach().cpu().clone() for k, v in model.model.state_dict().items()}
    pickled, log, epoch_preds = {}, [], {}

    def callback(epoch, stats):  # main.py:207-241: decode the training split and pickle the model every epoch
        epoch_preds[epoch] = model.predict(split)
        pickled[epoch] = pickle.dumps(model)
        log.append(stats['train_loss'])

    model.fit(split, use_labels=False, callback_fn=callback)
    assert len(log) == 2 and log[1] < log[0], log
    best = pickle.loads(pickled[int(np.argmin(log))])  # main.py:252-255
    assert best._make_data_loader is None and best.model.gaussian_means.is_cuda
    pred = best.predict(split)
    same, total = frames_equal(pred, epoch_preds[int(np.argmin(log))])
    assert same == total, "the unpickled model decodes differently from the live one"
    assert model._cache is not None and model._cache.hits > 0  # epoch 1 and the decodes were served from HBM

    # the reference wrapper from the same seed (init_logits is the only random draw; both wrappers initialise the
    # Gaussian from the same first shuffled batch): same losses after the same two epochs
    rargs = copy.copy(args)
    rargs.cuda = False
    torch.manual_seed(11)
    ref = ref_sm.SemiMarkovModel.from_args(rargs, split)
    assert torch.equal(ref.model.init_logits.detach(), init_state['init_logits'])
    ref_log = []
    ref.fit(split, use_labels=False, callback_fn=lambda epoch, stats: ref_log.append(float(stats['train_loss'])))
    assert np.allclose(log, ref_log, rtol=2e-4), (log, ref_log)

    # decode with the trained parameters in the reference's own module: identical predictions and metrics
    ref.model.load_state_dict({k: v.detach().cpu() for k, v in best.model.state_dict().items()})
    ref_pred = ref.predict(split)
    same, total = frames_equal(pred, ref_pred)
    assert total - same <= 0.002 * total, "%d of %d frames differ from the reference's decode" % (total - same, total)
    if same == total:
        assert_same_metrics(split, pred, ref_pred)


@needs_ref
def test_reference_main_train_and_test_drive_the_b200_classifier():
    """CLASSIFIERS['semimarkov'] = the B200 SemiMarkovModel, then the reference's own
    main.train (fit, per-epoch callback decode + pickle.dumps, pickle.loads of the best epoch) and main.test."""
    main = ref_import.load_reference_module('main')
    parser = argparse.ArgumentParser()
    main.add_serialization_args(parser)
    main.add_data_args(parser)
    old = main.CLASSIFIERS['semimarkov']
    main.CLASSIFIERS['semimarkov'] = pkg.SemiMarkovModel
    try:
        main.add_classifier_args(parser)
        main.add_training_args(parser)
        main.add_misc_args(parser)
        args = parser.parse_args(
            "--classifier semimarkov --training unsupervised --mix_tasks --task_specific_steps --sm_constrain_transitions "
            "--annotate_background_with_previous --sm_constrain_with_narration train --sm_constrain_narration_weight=-1e4 "
            "--cuda --epochs 2 --batch_size 4 --sm_max_span_length 12 --print_every 0".split())
        train = data.make_crosstask_like(n_tasks=2, steps_per_task=(2, 3), n_videos=10, feature_dim=8, frames=(40, 60),
                                         narration=True, seed=9)
        dev = data.make_crosstask_like(n_tasks=2, steps_per_task=(2, 3), n_videos=4, feature_dim=8, frames=(40, 60),
                                       narration=True, seed=9)
        torch.manual_seed(3)
        best = main.train(args, train, dev, 'split0', verbose=False)
        assert isinstance(best, pkg.SemiMarkovModel)
        stats = main.test(args, best, dev, 'dev', verbose=False)
        for task, st in stats.items():
            for key in main.STAT_KEYS:
                assert key in st, key
            assert 0.0 <= st['mof'][0] / st['mof'][1] <= 1.0
        # supervised S6 command line (README.md:43)
        args = parser.parse_args("--classifier semimarkov --training supervised --cuda".split())
        sup = data.make_supervised_like(n_tasks=2, n_videos=8, feature_dim=8, frames=(50, 80), seed=2)
        best = main.train(args, sup, sup, 'split0', verbose=False)
        stats = main.test(args, best, sup, 'test', verbose=False)
        mof = sum(st['mof'][0] for st in stats.values()) / sum(st['mof'][1] for st in stats.values())
        assert mof > 0.5, mof
    finally:
        main.CLASSIFIERS['semimarkov'] = old


def test_em_trainer_reachable_from_fit_and_improves_likelihood():
    split = data.make_crosstask_like(n_tasks=2, steps_per_task=(2, 3), n_videos=8, feature_dim=8, frames=(40, 60), seed=1)
    args = HsmmArgs(sm_max_span_length=12, sm_constrain_transitions=True, annotate_background_with_previous=True,
                    epochs=3, batch_size=4, training='unsupervised', sm_unsupervised_method='em')
    torch.manual_seed(0)
    model = pkg.SemiMarkovModel.from_args(args, split)
    log = []
    model.fit(split, use_labels=False, callback_fn=lambda e, s: log.append(s['train_loss']))
    assert len(log) == 3 and log[2] <= log[1] <= log[0] + 1e-6, log


def test_init_non_projection_parameters_from(tmp_path):
    split = data.make_supervised_like(n_tasks=1, n_videos=4, feature_dim=6, frames=(30, 40), seed=4)
    args = HsmmArgs(sm_max_span_length=8)
    model = pkg.SemiMarkovModel.from_args(args, split)
    model.fit(split, use_labels=True)
    path = tmp_path / "model.pkl"
    with open(path, "wb") as f:
        pickle.dump(model, f)
    args2 = HsmmArgs(sm_max_span_length=8, sm_init_non_projection_parameters_from=str(path), epochs=0, training='unsupervised')
    m2 = pkg.SemiMarkovModel.from_args(args2, split)
    for k, v in model.model.state_dict().items():
        assert torch.equal(m2.model.state_dict()[k].cpu(), v.cpu()), k
    calls = []
    m2.fit(split, use_labels=False, callback_fn=lambda e, s: calls.append(e))
    assert calls == [-1]  # semimarkov.py:172-175: no re-initialisation, one callback before training
    for k, v in model.model.state_dict().items():
        assert torch.equal(m2.model.state_dict()[k].cpu(), v.cpu()), k


# ---------------------------------------------------------------------------------------------
# module-level rows of SURVEY.md section 8(a) that had no test in round 1
# ---------------------------------------------------------------------------------------------
def test_initialize_gaussian_matches_reference_formula():
    """a13 (semimarkov_modules.py:263-282): every class mean = mean of the live frames, cov = diag of the UNBIASED
    variance (torch `var`), computed here by hsmm_feature_moments in f64."""
    torch.manual_seed(2)
    C, D, B, T = 6, 33, 5, 90
    m = pkg.SemiMarkovModule(HsmmArgs(sm_max_span_length=10), C, D, allow_self_transitions=True).cuda()
    data_ = torch.randn(B, T, D) * 3 + torch.arange(D) * 0.1
    lengths = torch.LongTensor([90, 41, 7, 90, 1])
    m.initialize_gaussian(data_.cuda(), lengths)
    feats = torch.cat([data_[i, :int(lengths[i])] for i in range(B)], dim=0).double()
    mean, var = feats.mean(dim=0), feats.var(dim=0)
    assert torch.allclose(m.gaussian_means.cpu().double(), mean.unsqueeze(0).expand(C, D), rtol=1e-6, atol=1e-6)
    assert torch.allclose(torch.diagonal(m.gaussian_cov).cpu().double(), var, rtol=1e-5)
    assert float((m.gaussian_cov - torch.diag(torch.diagonal(m.gaussian_cov))).abs().max()) == 0.0
    if ref_import.reference_available():  # and against the reference method itself, on its own module
        mods, _ = ref_import.load_reference()
        r = mods.SemiMarkovModule(HsmmArgs(sm_max_span_length=10), C, D, allow_self_transitions=True)
        r.initialize_gaussian(data_, lengths)
        assert torch.allclose(m.gaussian_means.cpu(), r.gaussian_means.detach(), rtol=1e-5, atol=1e-5)
        assert torch.allclose(m.gaussian_cov.cpu(), r.gaussian_cov.detach(), rtol=1e-4, atol=1e-6)


@needs_ref
def test_gold_score_ragged_batch_matches_reference_module():
    """a10 with videos of different lengths (the golden fixture has equal lengths): gold-path score and its gradients
    against the reference module's `log_likelihood(spans=...)` run on the CPU ONE VIDEO AT A TIME.

    Per video, because on a zero-padded ragged batch the reference's own number is not the model's score: `to_parts`
    (semimarkov_modules.py:641-642) ignores `lengths`, so the label-0 padding behind a shorter video's EOS is scored as
    extra class-0 segments, each EOS->0 edge costing -1e9 (a batch like this one gives ll ~ -1.5e9 there, and padding
    counts leak into the class-0 gradients).  The kernel scores the live frames only; DESIGN.md lists the deviation."""
    mods, utils = ref_import.load_reference()
    torch.manual_seed(4)
    C, D, K, B, T = 5, 7, 9, 4, 40
    args = HsmmArgs(sm_max_span_length=K)
    mine = pkg.SemiMarkovModule(args, C, D, allow_self_transitions=True).cuda()
    with torch.no_grad():
        mine.gaussian_means.normal_()
        mine.transition_logits.normal_()
        mine.poisson_log_rates.uniform_(0.5, 1.5)
    ref = mods.SemiMarkovModule(args, C, D, allow_self_transitions=True)
    ref.load_state_dict({k: v.cpu() for k, v in mine.state_dict().items()})
    lengths = torch.LongTensor([40, 23, 11, 31])
    labels = torch.randint(0, C, (B, 8)).repeat_interleave(5, dim=1)
    feats = torch.randn(B, T, D)
    for b in range(B):
        feats[b, int(lengths[b]):] = 0
        labels[b, int(lengths[b]):] = 0
    spans = utils.labels_to_spans(labels, max_k=K)
    for disc in (False, True):
        args.sm_train_discriminatively = disc
        mine.zero_grad()
        ref.zero_grad()
        ll, _ = mine.log_likelihood(feats.cuda(), lengths, None, spans=spans.cuda(), add_eos=True, use_mean_z=True)
        ll.backward()
        ll_r = 0.0
        for b in range(B):
            n = int(lengths[b])
            one, _ = ref.log_likelihood(feats[b:b + 1, :n], lengths[b:b + 1], None, spans=spans[b:b + 1, :n], add_eos=True,
                                        use_mean_z=True)
            (one / B).backward()
            ll_r += float(one) / B
        assert abs(float(ll) - ll_r) <= 1e-4 * abs(ll_r), (disc, float(ll), ll_r)
        for k in ("gaussian_means", "transition_logits", "init_logits", "poisson_log_rates"):
            a, b_ = getattr(mine, k).grad.cpu().numpy(), getattr(ref, k).grad.numpy()
            assert np.abs(a - b_).max() <= 1e-4 * max(1e-12, np.abs(b_).max()), (disc, k)
    args.sm_train_discriminatively = False


def test_gold_score_rejects_unscorable_segmentations():
    """Labels outside the batch's valid classes and gold segments longer than K-1 make the reference raise
    (struct.to_parts / score, semimarkov_modules.py:626-655); here the video's score is NaN -- never a silent number."""
    torch.manual_seed(0)
    C, D, K = 6, 4, 5
    m = pkg.SemiMarkovModule(HsmmArgs(sm_max_span_length=K), C, D, allow_self_transitions=True).cuda()
    feats = torch.randn(2, 12, D).cuda()
    lengths = torch.LongTensor([12, 12])
    valid = [torch.LongTensor([0, 2, 3]), torch.LongTensor([0, 2, 3])]
    ok = torch.LongTensor([[0, -1, 2, -1, -1, 3, -1, 0, -1, -1, 2, -1]] * 2)
    ll, _ = m.log_likelihood(feats, lengths, valid, spans=ok.cuda())
    assert torch.isfinite(ll)
    bad_label = ok.clone()
    bad_label[1, 5] = 4  # class 4 is not valid for this batch
    ll, _ = m.log_likelihood(feats, lengths, valid, spans=bad_label.cuda())
    assert torch.isnan(ll)
    too_long = ok.clone()
    too_long[0] = torch.LongTensor([0, -1, -1, -1, -1, -1, 2, -1, 0, -1, -1, 3])  # a 6-frame segment, K-1 = 4
    ll, _ = m.log_likelihood(feats, lengths, valid, spans=too_long.cuda())
    assert torch.isnan(ll)


def test_combined_loglik_and_viterbi_equals_separate_calls():
    """log_likelihood_and_viterbi scores the emissions once; results and gradients equal the two separate calls."""
    split = data.make_crosstask_like(n_tasks=1, steps_per_task=(4, 4), n_videos=4, feature_dim=12, frames=(50, 80), narration=True, seed=8)
    args = HsmmArgs(sm_max_span_length=15, sm_constrain_transitions=True, annotate_background_with_previous=True,
                    sm_constrain_with_narration=['train'], batch_size=4)
    torch.manual_seed(1)
    model = pkg.SemiMarkovModel.from_args(args, split)
    with torch.no_grad():
        model.model.gaussian_means.normal_()
    batch = next(iter(model._device_batches(split, model._loader(split, shuffle=False, batch_by_task=True, batch_size=4))))
    pen = model._narration(split, batch, 'train')
    addl = model.make_additional_allowed_ends(batch['task_name'], batch['lengths'])
    m = model.model
    ll, _, spans, labels = m.log_likelihood_and_viterbi(batch['features'], batch['lengths'], batch['task_indices'],
                                                        additional_allowed_ends_per_instance=addl, constraints=pen)
    ll.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    ll2, _ = m.log_likelihood(batch['features'], batch['lengths'], batch['task_indices'],
                              additional_allowed_ends_per_instance=addl, constraints=pen)
    ll2.backward()
    spans2, labels2 = m.viterbi(batch['features'], batch['lengths'], batch['task_indices'],
                                additional_allowed_ends_per_instance=addl, constraints=pen, return_labels=True)
    assert float(ll) == float(ll2)
    assert torch.equal(spans, spans2) and torch.equal(labels, labels2)
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, grads[k], rtol=1e-4, atol=1e-5), k  # atomics: summation order varies


def test_viterbi_batches_grouped_equals_per_batch_viterbi():
    """SemiMarkovModule.viterbi_batches (one grouped launch for up to 32 mini-batches of different tasks) against
    `viterbi` called batch by batch; a second model without transition constraints takes the per-batch route."""
    split = data.make_crosstask_like(n_tasks=4, steps_per_task=(2, 6), n_videos=37, feature_dim=10, frames=(30, 90), narration=True,
                                     seed=21, allow_short=True)
    for constrained in (True, False):
        args = HsmmArgs(sm_max_span_length=15, sm_constrain_transitions=constrained, annotate_background_with_previous=True,
                        sm_constrain_with_narration=['test'], batch_size=3)
        torch.manual_seed(7)
        model = pkg.SemiMarkovModel.from_args(args, split)
        with torch.no_grad():
            model.model.gaussian_means.normal_()
            model.model.transition_logits.normal_()
            model.model.poisson_log_rates.uniform_(0.5, 2.0)
        loader = model._loader(split, shuffle=False, batch_by_task=True, batch_size=3)
        batches = []
        for batch in model._device_batches(split, loader):
            batches.append(dict(features=batch['features'], lengths=batch['lengths'], valid_classes_per_instance=batch['task_indices'],
                                additional_allowed_ends_per_instance=model.make_additional_allowed_ends(batch['task_name'], batch['lengths']),
                                constraints=model._narration(split, batch, 'test')))
        assert len(batches) > 8
        grouped = model.model.viterbi_batches(batches, non_blocking=False)
        for bt, (spans, labels) in zip(batches, grouped):
            ref_spans, ref_labels = model.model.viterbi(bt['features'], bt['lengths'], bt['valid_classes_per_instance'],
                                                        additional_allowed_ends_per_instance=bt['additional_allowed_ends_per_instance'],
                                                        constraints=bt['constraints'], return_labels=True)
            assert torch.equal(spans, ref_spans) and torch.equal(labels, ref_labels)
        pred = model.predict(split)
        assert len(pred) == 37 and all(len(pred[v['video_name']]) == v['features'].shape[0] for v in split.videos)
