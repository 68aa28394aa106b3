"""`SemiMarkovModel`: the classifier/trainer wrapper the reference's CLI registers as
`CLASSIFIERS['semimarkov']` (/root/reference/src/main.py:36-45, models/semimarkov/semimarkov.py).

Same flags, `from_args`, `fit(train_data, use_labels, callback_fn)` and `predict(test_data)`
contracts; the DP work goes through the B200 `SemiMarkovModule`.  Data access is by duck typing on
the reference's `Datasplit` API (`corpus.n_classes`, `feature_dim`, `get_allowed_starts_and_transitions`,
`get_ordered_indices_no_background`, and a loader that yields `padding_colate` batch dicts,
models/model.py:42-63), so the reference's own data layer plugs in unchanged; `data.py` provides a
synthetic stand-in with the same surface.

What differs from the reference's wrapper, by design:
  * batches live on the device after their first use (`DeviceBatchCache`): the reference's BatchSampler
    (data/corpus.py:613-644) always forms the same batches, so from the second epoch -- and in the
    per-epoch decode of the training split (main.py:207-218) -- nothing is read, padded or copied again;
  * `predict` enqueues every batch's Viterbi kernels and result copies (pinned memory) and synchronises
    ONCE per split instead of once per batch; per-frame labels come from the kernel in global ids;
  * multi-GPU: every rank takes its shard of each mini-batch and the packed gradient buffer is
    all-reduced once per optimiser step (distributed.py); parameters are broadcast from rank 0.
The model object pickles (main.py:234, 248-257, 445-469): process groups, loaders and device caches are
dropped from the pickled state and rebuilt lazily.
"""
import pickle
import time

import numpy as np
import torch

from . import distributed as hdist
from . import semimarkov_utils
from .semimarkov_modules import SemiMarkovModule, all_equal


def make_optimizer(args, parameters):
    """models/model.py:27-39 (Adam + ReduceLROnPlateau; `verbose` no longer exists in torch 2.x)."""
    opt = torch.optim.Adam(parameters, lr=args.lr)
    scheduler = None
    if not getattr(args, 'no_reduce_plateau', False):
        scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(
            opt, factor=args.reduce_plateau_factor, patience=int(args.reduce_plateau_patience), min_lr=1e-4, threshold=1e-5)
    return opt, scheduler


def resolve_data_loader():
    """Inside the reference tree: its own loader (models/model.py:66-77); otherwise the stand-in."""
    try:
        from models.model import make_data_loader
    except ImportError:
        from .data import make_data_loader
    return make_data_loader


class DeviceBatchCache:
    """Batches of a split, resident in HBM after their first use.  Key = the sampler's batch (a tuple of
    (task, video) keys, data/corpus.py:633-636), so a hit skips `Datasplit.__getitem__`, `padding_colate`
    and the host->device copy altogether.  CrossTask's PCA features are ~4.4 GB; a B200 has 180 GB."""

    MOVE = ('features', 'gt_single', 'constraints')

    def __init__(self, limit_bytes=64 << 30):
        self.entries, self.bytes, self.limit = {}, 0, limit_bytes
        self.hits = self.misses = 0

    def get(self, key, build):
        e = self.entries.get(key)
        if e is not None:
            self.hits += 1
            return e
        self.misses += 1
        e = self.to_device(build())
        size = sum(v.numel() * v.element_size() for v in e.values() if isinstance(v, torch.Tensor) and v.is_cuda)
        if self.bytes + size <= self.limit:
            self.entries[key] = e
            self.bytes += size
        return e

    @classmethod
    def to_device(cls, batch):
        out = dict(batch)
        for k in cls.MOVE:
            if k in out and isinstance(out[k], torch.Tensor) and not out[k].is_cuda:
                src = out[k]
                if not src.is_pinned():
                    try:
                        src = src.pin_memory()
                    except RuntimeError:
                        pass
                out[k] = src.cuda(non_blocking=True)
        return out


class SemiMarkovModel(object):
    @classmethod
    def add_args(cls, parser):
        # models/semimarkov/semimarkov.py:16-31
        SemiMarkovModule.add_args(parser)
        parser.add_argument('--sm_component_model', action='store_true')
        parser.add_argument('--sm_constrain_transitions', action='store_true')
        parser.add_argument('--sm_constrain_with_narration', choices=['train', 'test'], nargs='*', default=[])
        parser.add_argument('--sm_constrain_narration_weight', type=float, default=-1e4)
        parser.add_argument('--sm_train_discriminatively', action='store_true')
        parser.add_argument('--sm_hidden_markov', action='store_true')
        parser.add_argument('--sm_predict_single', action='store_true')
        # extensions: closed-form EM beside the reference's Adam-on-logZ; device cache switch
        parser.add_argument('--sm_unsupervised_method', choices=['gradient', 'em'], default='gradient')
        parser.add_argument('--sm_no_device_cache', action='store_true')

    @classmethod
    def from_args(cls, args, train_data, make_data_loader=None):
        # models/semimarkov/semimarkov.py:33-114
        n_classes = train_data.corpus.n_classes
        feature_dim = train_data.feature_dim
        allow_self_transitions = True
        assert args.sm_max_span_length is not None
        if getattr(args, 'sm_component_model', False):
            raise NotImplementedError("--sm_component_model is outside the B200 hot path (SURVEY.md section 2 row 7)")
        if args.sm_constrain_transitions:
            allowed_starts, allowed_transitions, allowed_ends, ordered_indices_by_task = \
                train_data.get_allowed_starts_and_transitions()
            for src in range(n_classes):
                allowed_transitions.setdefault(src, set()).add(src)
        else:
            allowed_starts = allowed_transitions = allowed_ends = ordered_indices_by_task = None
        merge_classes = None
        if getattr(args, 'annotate_background_with_previous', False) and not getattr(args, 'no_merge_classes', False):
            merge_classes = {}
            bkg = set(train_data.corpus._background_indices)
            for task, indices in train_data.corpus._indices_by_task.items():
                background = [ix for ix in indices if ix in bkg]
                canon = background[0]
                for ix in indices:
                    tgt = canon if ix in bkg else ix
                    assert merge_classes.setdefault(ix, tgt) == tgt
        model = SemiMarkovModule(args, n_classes, feature_dim, allow_self_transitions=allow_self_transitions,
                                 allowed_starts=allowed_starts, allowed_transitions=allowed_transitions,
                                 allowed_ends=allowed_ends, merge_classes=merge_classes)
        return cls(args, n_classes, feature_dim, model, ordered_indices_by_task, make_data_loader=make_data_loader)

    def __init__(self, args, n_classes, feature_dim, model, ordered_indices_by_task=None, make_data_loader=None,
                 dist_group=None):
        self.args = args
        self.n_classes = n_classes
        self.feature_dim = feature_dim
        self.model = model
        self.ordered_indices_by_task = ordered_indices_by_task
        self.dist_group = dist_group
        self._make_data_loader = make_data_loader  # None: resolved on every use (never pickled)
        self._cache = None
        if torch.cuda.is_available():  # without a GPU the first kernel call raises HsmmError (there is no CPU path)
            self.model.cuda()

    # -- pickling (main.py:234 pickles the model every epoch; :248-257 and :445-469 load it back) -----
    def __getstate__(self):
        d = dict(self.__dict__)
        d['dist_group'] = None
        d['_make_data_loader'] = None
        d['_cache'] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.__dict__.setdefault('dist_group', None)
        self.__dict__.setdefault('_make_data_loader', None)
        self.__dict__.setdefault('_cache', None)
        if torch.cuda.is_available():
            self.model.cuda()

    # -- data plumbing --------------------------------------------------------------------------------
    def _loader(self, datasplit, **kw):
        fn = self._make_data_loader or resolve_data_loader()
        return fn(self.args, datasplit, **kw)

    def _device_batches(self, datasplit, loader):
        """Iterate `loader`, yielding batch dicts whose tensors are on the device.  With a batch sampler and a
        map-style dataset (the reference's DataLoader, data.py's stand-in) batches are served from the device
        cache after their first use."""
        sampler = getattr(loader, 'batch_sampler', None)
        dataset = getattr(loader, 'dataset', None)
        collate = getattr(loader, 'collate_fn', None)
        if getattr(self.args, 'sm_no_device_cache', False) or sampler is None or dataset is None or collate is None:
            for batch in loader:
                yield DeviceBatchCache.to_device(batch)
            return
        if self._cache is None:
            self._cache = DeviceBatchCache()
        for keys in sampler:
            keys = list(keys)
            ck = (id(datasplit), tuple(keys))
            yield self._cache.get(ck, lambda: collate([dataset[k] for k in keys]))

    def fit_supervised(self, train_data):
        # models/semimarkov/semimarkov.py:125-133
        assert not self.args.sm_constrain_transitions
        loader = self._loader(train_data, batch_by_task=False, shuffle=False, batch_size=1)
        features, labels = [], []
        for batch in loader:
            features.append(batch['features'].squeeze(0))
            labels.append(batch['gt_single'].squeeze(0))
        self.model.fit_supervised(features, labels)

    def make_additional_allowed_ends(self, tasks, lengths):
        # models/semimarkov/semimarkov.py:135-147: a video shorter than its task chain may end early
        if self.ordered_indices_by_task is None:
            return None
        out = []
        for task, length in zip(tasks, lengths):
            ord_indices = self.ordered_indices_by_task[task]
            n = int(length)
            out.append([ord_indices[n - 1]] if n < len(ord_indices) else [])
        return out

    def expand_constraints(self, datasplit, task, task_indices, constraints):
        # models/semimarkov/semimarkov.py:149-157: (B, T, n_steps) -> (B, T, C) in the step columns
        task_indices = [int(x) for x in task_indices.cpu()]
        step_indices = datasplit.get_ordered_indices_no_background()[task]
        assert constraints.size(2) == len(step_indices)
        expanded = torch.zeros((constraints.size(0), constraints.size(1), len(task_indices)), device=constraints.device)
        cols = torch.as_tensor([task_indices.index(label) for label in step_indices], dtype=torch.long,
                               device=constraints.device)
        expanded[:, :, cols] = constraints
        return expanded

    def _narration(self, datasplit, batch, which):
        """Additive narration penalty (B, T, C) of models/semimarkov/semimarkov.py:227-234, 341-348; built on the device
        and kept in the (cached) batch dict."""
        if which not in self.args.sm_constrain_with_narration:
            return None
        if '_narration_penalty' not in batch:
            tasks = batch['task_name']
            assert all_equal(tasks)
            c = self.expand_constraints(datasplit, tasks[0], batch['task_indices'][0], 1 - batch['constraints'])
            batch['_narration_penalty'] = (c * self.args.sm_constrain_narration_weight).cuda()
        return batch['_narration_penalty']

    # -- training -------------------------------------------------------------------------------
    def fit(self, train_data, use_labels, callback_fn=None):
        # models/semimarkov/semimarkov.py:159-316
        args = self.args
        self.model.train()
        if use_labels:
            assert not args.sm_constrain_transitions
        initialize = True
        if use_labels and args.sm_supervised_method in ['closed-form', 'closed-then-gradient']:
            self.fit_supervised(train_data)
            if args.sm_supervised_method == 'closed-then-gradient':
                initialize = False
                if callback_fn:
                    callback_fn(-1, {})
            else:
                return
        if getattr(args, 'sm_init_non_projection_parameters_from', None):
            initialize = False  # parameters came from the pickled model (SemiMarkovModule.__init__)
            if callback_fn:
                callback_fn(-1, {})
        optimizer, scheduler = make_optimizer(args, self.model.parameters())
        if initialize:
            big = next(iter(self._loader(train_data, batch_by_task=False, shuffle=True, batch_size=100)))
            self.model.initialize_gaussian(big['features'].cuda(), big['lengths'])
        rank, world = hdist.rank_world(self.dist_group)
        # replicas must start identical: init_logits is drawn per process (semimarkov_modules.py:159)
        hdist.broadcast_parameters(self.model, self.dist_group)
        if not use_labels and getattr(args, 'sm_unsupervised_method', 'gradient') == 'em':
            return self._fit_em(train_data, callback_fn)
        loader = self._loader(train_data, batch_by_task=True, shuffle=True, batch_size=args.batch_size)
        K = args.sm_max_span_length
        dev = self.model.gaussian_means.device
        for epoch in range(args.epochs):
            start_time = time.time()
            self.model.train()
            losses, pending = [], []
            weights = []
            num_frames = num_videos = 0
            for batch_ix, batch in enumerate(self._device_batches(train_data, loader)):
                if getattr(args, 'train_limit', None) and batch_ix >= args.train_limit:
                    break
                tasks, lengths = batch['task_name'], batch['lengths']
                constraints = self._narration(train_data, batch, 'train')
                num_frames += int(lengths.sum())
                num_videos += len(lengths)
                addl = self.make_additional_allowed_ends(tasks, lengths)
                # data-parallel shard of the mini-batch (videos are independent given the parameters); a rank
                # without videos (batch smaller than the world) contributes zero gradients and still joins the all-reduce
                sel = hdist.shard_indices(len(lengths), rank, world)
                if len(sel) > 0:
                    whole = len(sel) == len(lengths)
                    features = batch['features'] if whole else batch['features'][sel]
                    sub_lengths = lengths if whole else lengths[sel]
                    spans = None
                    if use_labels:
                        gt = batch['gt_single'] if whole else batch['gt_single'][sel]
                        spans = semimarkov_utils.labels_to_spans(gt, max_k=K)
                    ll, log_det = self.model.log_likelihood(
                        features, sub_lengths, valid_classes_per_instance=[batch['task_indices'][i] for i in sel],
                        spans=spans, add_eos=True, use_mean_z=use_labels,
                        additional_allowed_ends_per_instance=None if addl is None else [addl[i] for i in sel],
                        constraints=None if constraints is None else (constraints if whole else constraints[sel]))
                    # `ll` is the mean over this rank's shard; weight it so that the all-reduced SUM of
                    # gradients equals the gradient of the mean over the whole mini-batch
                    frac = len(sel) / float(len(lengths))
                    pending.append(-(ll * frac) - log_det * frac)
                else:
                    pending.append(None)
                if len(pending) >= args.batch_accumulation:
                    live = [p for p in pending if p is not None]
                    n_pending = len(pending)
                    pending = []
                    if live:
                        loss = sum(live) / n_pending
                        loss.backward()
                        loss_val = loss.detach()
                    else:
                        loss_val = torch.zeros((), device=dev)
                    loss_val = hdist.allreduce_gradients(self.model.parameters(), loss_val, self.dist_group)
                    losses.append(loss_val.reshape(()))  # stays on the device: one synchronisation per epoch
                    weights.append(len(lengths))
                    if args.max_grad_norm is not None:
                        torch.nn.utils.clip_grad_norm_(self.model.parameters(), args.max_grad_norm)
                    optimizer.step()
                    self.model.zero_grad()
                    if getattr(args, 'print_every', 0) and batch_ix % args.print_every == 0 and rank == 0:
                        nll_so_far = float((torch.stack(losses) * torch.tensor(weights, device=dev)).sum())
                        print('Epoch: %02d, Batch: %03d, loss: %.4f, recon: %.4f, Throughput: %.2f vid / sec' % (
                            epoch, batch_ix, nll_so_far / max(num_videos, 1), nll_so_far / max(num_frames, 1),
                            num_videos / (time.time() - start_time)))
            if losses:
                lv = torch.stack(losses).double()
                train_loss = float(lv.mean())
                train_nll = float((lv * torch.tensor(weights, device=dev, dtype=torch.float64)).sum())
            else:
                train_loss, train_nll = float('nan'), 0.0
            if scheduler is not None:
                scheduler.step(train_loss)
            if callback_fn:
                callback_fn(epoch, {'train_loss': train_loss,
                                    'train_nll_frame_avg': train_nll / max(num_frames, 1),
                                    'train_kl_vid_avg': 0.0,
                                    'train_recon_bound': train_nll / max(num_frames, 1)})

    def _fit_em(self, train_data, callback_fn=None):
        """Closed-form EM (--sm_unsupervised_method em): one E-step over the split with
        the same forward/backward kernels, ONE all-reduce of the packed expected counts, closed-form M-step."""
        args = self.args
        loader = self._loader(train_data, batch_by_task=True, shuffle=False, batch_size=args.batch_size)
        rank, world = hdist.rank_world(self.dist_group)
        for epoch in range(args.epochs):
            self.model.train()
            total = None
            num_frames = 0
            for batch_ix, batch in enumerate(self._device_batches(train_data, loader)):
                if getattr(args, 'train_limit', None) and batch_ix >= args.train_limit:
                    break
                num_frames += int(batch['lengths'].sum())
                if batch_ix % world != rank:  # whole batches per rank: the statistics simply add up
                    continue
                constraints = self._narration(train_data, batch, 'train')
                addl = self.make_additional_allowed_ends(batch['task_name'], batch['lengths'])
                st = self.model.expected_statistics(batch['features'], batch['lengths'], batch['task_indices'],
                                                    additional_allowed_ends_per_instance=addl, constraints=constraints)
                total = self.model.add_statistics(total, st)
            buf = self.model.pack_statistics(total) if total is not None else \
                torch.zeros(self.model.statistics_size(), device=self.model.gaussian_means.device)
            buf = hdist.allreduce_stats(buf, self.dist_group)
            stats = self.model.unpack_statistics(buf)
            ll = self.model.em_update(stats)
            if callback_fn:
                nll = -float(stats['logz'])
                callback_fn(epoch, {'train_loss': -ll, 'train_nll_frame_avg': nll / max(num_frames, 1),
                                    'train_kl_vid_avg': 0.0, 'train_recon_bound': nll / max(num_frames, 1)})

    # -- decoding -------------------------------------------------------------------------------
    def predict(self, test_data):
        """models/semimarkov/semimarkov.py:318-410.  Per-frame labels (global class ids) come straight from the
        Viterbi kernel; up to 32 mini-batches are decoded by one grouped launch (SemiMarkovModule.viterbi_batches), the
        result copies are enqueued back to back and the host synchronises once for the whole split."""
        self.model.eval()
        predictions = {}
        loader = self._loader(test_data, shuffle=False, batch_by_task=True, batch_size=self.args.batch_size)
        queued, pending = [], []

        def flush():
            res = self.model.viterbi_batches([p[0] for p in pending], return_labels=True, return_spans=False, non_blocking=True)
            for (_, labels), (_, videos, lengths) in zip(res, pending):
                queued.append((videos, labels, lengths))
            pending.clear()

        for batch in self._device_batches(test_data, loader):
            tasks, lengths = batch['task_name'], batch['lengths']
            assert len(set(tasks)) == 1
            pending.append((dict(features=batch['features'], lengths=lengths, valid_classes_per_instance=batch['task_indices'],
                                 additional_allowed_ends_per_instance=self.make_additional_allowed_ends(tasks, lengths),
                                 constraints=self._narration(test_data, batch, 'test')), batch['video_name'], lengths))
            if len(pending) >= self.model.GROUP_MAX:
                flush()
        if pending:
            flush()
        torch.cuda.current_stream().synchronize()  # all label copies have landed in pinned memory
        for videos, labels, lengths in queued:
            labels = labels.numpy()
            for video, lab, n in zip(videos, labels, lengths):
                preds = lab[:int(n)].copy()
                assert self.model.n_classes not in preds, "predictions should not contain EOS: {}".format(preds)
                predictions[video] = preds
        return predictions


def load_pickled_model(path):
    """utils/utils.py:load_pickle for a pickled SemiMarkovModel (main.py:445-469)."""
    with open(path, 'rb') as f:
        return pickle.load(f)
