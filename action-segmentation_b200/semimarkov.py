"""`SemiMarkovModel`: the classifier/trainer wrapper the reference's CLI registers as
`CLASSIFIERS['semimarkov']` (/root/reference/src/main.py:36-45, models/semimarkov/semimarkov.py).

Same flags, `from_args`, `fit(train_data, use_labels, callback_fn)` and `predict(test_data)`
contracts; the DP work goes through the B200 `SemiMarkovModule`.  Data access is by duck typing on
the reference's `Datasplit` API (`corpus.n_classes`, `feature_dim`, `get_allowed_starts_and_transitions`,
`get_ordered_indices_no_background`, and a loader that yields `padding_colate` batch dicts,
models/model.py:42-63), so the reference's own data layer plugs in unchanged; `data.py` provides a
synthetic stand-in with the same surface.

Multi-GPU: pass `dist_group` (or initialise torch.distributed) and every rank processes its shard
of each mini-batch; the packed gradient buffer is all-reduced once per optimiser step
(distributed.py).
"""
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
        if make_data_loader is None:
            try:  # inside the reference tree: its own loader (models/model.py:66-77)
                from models.model import make_data_loader
            except ImportError:
                from .data import make_data_loader
        self._make_data_loader = make_data_loader
        self.model.cuda()

    def __getstate__(self):
        # picklable like the reference's model object (main.py:234): drop the process group / loader fn
        d = dict(self.__dict__)
        d['dist_group'] = None
        d['_make_data_loader'] = None
        return d

    # -- helpers shared by fit / predict --------------------------------------------------------
    def fit_supervised(self, train_data):
        # models/semimarkov/semimarkov.py:125-133
        assert not self.args.sm_constrain_transitions
        loader = self._make_data_loader(self.args, train_data, batch_by_task=False, shuffle=False, batch_size=1)
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
        expanded = torch.zeros((constraints.size(0), constraints.size(1), len(task_indices)))
        cols = torch.as_tensor([task_indices.index(label) for label in step_indices], dtype=torch.long)
        expanded[:, :, cols] = constraints
        return expanded

    def _narration(self, datasplit, batch, which):
        if which not in self.args.sm_constrain_with_narration:
            return None
        tasks = batch['task_name']
        assert all_equal(tasks)
        c = self.expand_constraints(datasplit, tasks[0], batch['task_indices'][0], 1 - batch['constraints'])
        return (c * self.args.sm_constrain_narration_weight).cuda(non_blocking=True)

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
        optimizer, scheduler = make_optimizer(args, self.model.parameters())
        if initialize:
            big = next(iter(self._make_data_loader(args, train_data, batch_by_task=False, shuffle=True, batch_size=100)))
            self.model.initialize_gaussian(big['features'].cuda(), big['lengths'])
        loader = self._make_data_loader(args, train_data, batch_by_task=True, shuffle=True, batch_size=args.batch_size)
        K = args.sm_max_span_length
        rank, world = hdist.rank_world(self.dist_group)
        for epoch in range(args.epochs):
            start_time = time.time()
            self.model.train()
            losses, pending = [], []
            num_frames = num_videos = 0
            train_nll = 0.0
            for batch_ix, batch in enumerate(loader):
                if getattr(args, 'train_limit', None) and batch_ix >= args.train_limit:
                    break
                tasks, lengths = batch['task_name'], batch['lengths']
                constraints = self._narration(train_data, batch, 'train')
                num_frames += int(lengths.sum())
                num_videos += len(lengths)
                addl = self.make_additional_allowed_ends(tasks, lengths)
                # data-parallel shard of the mini-batch (videos are independent given the parameters)
                sel = hdist.shard_indices(len(lengths), rank, world)
                features = batch['features'][sel].cuda(non_blocking=True)
                sub_lengths = lengths[sel]
                spans = None
                if use_labels:
                    spans = semimarkov_utils.labels_to_spans(batch['gt_single'][sel].cuda(), max_k=K)
                ll, log_det = self.model.log_likelihood(
                    features, sub_lengths, valid_classes_per_instance=[batch['task_indices'][i] for i in sel],
                    spans=spans, add_eos=True, use_mean_z=use_labels,
                    additional_allowed_ends_per_instance=None if addl is None else [addl[i] for i in sel],
                    constraints=None if constraints is None else constraints[sel])
                # `ll` is the mean over this rank's shard; weight it so that the all-reduced SUM of
                # gradients equals the gradient of the mean over the whole mini-batch
                this_loss = -(ll * (len(sel) / float(len(lengths)))) - log_det
                pending.append(this_loss)
                if len(pending) >= args.batch_accumulation:
                    loss = sum(pending) / len(pending)
                    loss.backward()
                    pending = []
                    loss_val = hdist.allreduce_gradients(self.model.parameters(), loss.detach(), self.dist_group)
                    nll = float(loss_val)
                    losses.append(nll)
                    train_nll += nll * len(lengths)
                    if args.max_grad_norm is not None:
                        torch.nn.utils.clip_grad_norm_(self.model.parameters(), args.max_grad_norm)
                    optimizer.step()
                    self.model.zero_grad()
                    if getattr(args, 'print_every', 0) and batch_ix % args.print_every == 0 and rank == 0:
                        print('Epoch: %02d, Batch: %03d, loss: %.4f, recon: %.4f, Throughput: %.2f vid / sec' % (
                            epoch, batch_ix, train_nll / max(num_videos, 1), train_nll / max(num_frames, 1),
                            num_videos / (time.time() - start_time)))
            train_loss = float(np.mean(losses)) if losses else float('nan')
            if scheduler is not None:
                scheduler.step(train_loss)
            if callback_fn:
                callback_fn(epoch, {'train_loss': train_loss,
                                    'train_nll_frame_avg': train_nll / max(num_frames, 1),
                                    'train_kl_vid_avg': 0.0,
                                    'train_recon_bound': train_nll / max(num_frames, 1)})

    # -- decoding -------------------------------------------------------------------------------
    def predict(self, test_data):
        # models/semimarkov/semimarkov.py:318-410; per-frame labels come straight from the kernel
        self.model.eval()
        predictions = {}
        loader = self._make_data_loader(self.args, test_data, shuffle=False, batch_by_task=True,
                                        batch_size=self.args.batch_size)
        for batch in loader:
            tasks, lengths = batch['task_name'], batch['lengths']
            assert len(set(tasks)) == 1
            constraints = self._narration(test_data, batch, 'test')
            addl = self.make_additional_allowed_ends(tasks, lengths)
            _, labels = self.model.viterbi(batch['features'].cuda(non_blocking=True), lengths, batch['task_indices'],
                                           add_eos=True, use_mean_z=True, additional_allowed_ends_per_instance=addl,
                                           constraints=constraints, return_labels=True)
            for video, lab, n in zip(batch['video_name'], labels, lengths):
                preds = lab[:int(n)].numpy()
                assert self.model.n_classes not in preds, "predictions should not contain EOS: {}".format(preds)
                predictions[video] = preds
        return predictions
