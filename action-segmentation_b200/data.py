"""Synthetic stand-in for the reference's data layer, exposing exactly the surface the HSMM wrapper
touches (src/data/corpus.py: Datasplit.__getitem__ :315-380, BatchSampler :613-644; src/data/crosstask.py:
get_allowed_starts_and_transitions :328-388; src/models/model.py: padding_colate :42-63,
make_data_loader :66-77).  The reference's datasets are not available offline; benchmarks and tests
use CrossTask-/Breakfast-shaped synthetic videos instead."""
import random

import numpy as np
import torch


class SyntheticCorpus:
    def __init__(self, n_classes, indices_by_task, background_indices, annotate_background_with_previous=False):
        self.n_classes = n_classes
        self._indices_by_task = indices_by_task
        self._background_indices = list(background_indices)
        self.index2label = {i: ('BKG_%d' % i if i in set(background_indices) else 'step_%d' % i) for i in range(n_classes)}
        self.annotate_background_with_previous = annotate_background_with_previous

    def indices_by_task(self, task):
        return self._indices_by_task[task]


class SyntheticDatasplit:
    """videos: list of dicts with task_name, video_name, features (T,D), gt_single (T), constraints (T,n_steps) or None."""

    def __init__(self, corpus, videos, feature_dim, chains=None, remove_background=False):
        self.corpus = corpus
        self.videos = videos
        self.feature_dim = feature_dim
        self.chains = chains  # task -> ordered class ids bkg, step, bkg, ...
        self.remove_background = remove_background

    def __len__(self):
        return len(self.videos)

    def __getitem__(self, i):
        v = self.videos[i]
        d = {k: v[k] for k in ('task_name', 'video_name', 'features', 'gt_single')}
        d['task_indices'] = torch.LongTensor(sorted(self.corpus.indices_by_task(v['task_name'])))
        if v.get('constraints') is not None:
            d['constraints'] = v['constraints']
        return d

    def get_ordered_indices_no_background(self):
        bkg = set(self.corpus._background_indices)
        return {t: [c for c in ch if c not in bkg] for t, ch in self.chains.items()}

    def get_allowed_starts_and_transitions(self):
        allowed_starts, allowed_ends, allowed_transitions = set(), set(), {}
        for ch in self.chains.values():
            for s, t in zip(ch, ch[1:]):
                allowed_transitions.setdefault(s, set()).add(t)
            allowed_starts.add(ch[0])
            allowed_ends.add(ch[-1])
        return allowed_starts, allowed_transitions, allowed_ends, dict(self.chains)

    # -- the rest of the reference's Datasplit surface that main.py / the reference wrapper touch -------------
    @property
    def _corpus(self):
        return self.corpus

    def batch_sampler(self, batch_size=1, batch_by_task=True, shuffle=False):
        """data/corpus.py:300-301; lets the reference's own torch DataLoader (models/model.py:66-77) run on this split."""
        return self.batches(batch_size, batch_by_task, shuffle)

    def canonicalize_background(self, index):
        # data/corpus.py:398-402
        return self.corpus._background_indices[0] if index in self.corpus._background_indices else index

    def accuracy_corpus(self, optimal_assignment, prediction_function, prefix='', verbose=True, compare_to_folder=None):
        """data/corpus.py:405-604 with the vectorised metrics of evaluation.py: {task: {statistic: [num, den]}}."""
        from .evaluation import segmentation_metrics

        class _V:
            def __init__(self, name):
                self.name = name

        by_task = {}
        for v in self.videos:
            by_task.setdefault(v['task_name'], []).append(v)
        stats_by_task = {}
        canon = self.corpus.annotate_background_with_previous
        for task, vids in by_task.items():
            gts, preds = [], []
            for v in sorted(vids, key=lambda x: x['video_name']):
                gt = [int(x) for x in v['gt_single']]
                pred = [int(x) for x in prediction_function(_V(v['video_name']))]
                if canon:
                    gt = [self.canonicalize_background(x) for x in gt]
                    pred = [self.canonicalize_background(x) for x in pred]
                gts.append([[x] for x in gt])
                preds.append(np.asarray(pred))
            st = segmentation_metrics(gts, preds, self.corpus._background_indices, optimal_assignment=optimal_assignment)
            st['num_videos'] = np.array([len(vids), 1])
            stats_by_task[task] = st
        return stats_by_task

    def batches(self, batch_size, batch_by_task, shuffle, seed=1):
        """data/corpus.py:613-644 (BatchSampler): batches are ALWAYS task-homogeneous slices of the sorted video
        names of each task (`batch_by_task` is ignored there too); `shuffle` permutes the order of the batches
        with a fixed seed, never their composition."""
        by_task = {}
        for i, v in enumerate(self.videos):
            by_task.setdefault(v['task_name'], []).append(i)
        out = []
        for t in sorted(by_task):
            ids = sorted(by_task[t], key=lambda i: self.videos[i]['video_name'])
            out.extend(ids[i:i + batch_size] for i in range(0, len(ids), batch_size))
        if shuffle:
            random.Random(seed).shuffle(out)
        return out


def padding_colate(samples):
    samples = [s for s in samples if s is not None]
    keys = samples[0].keys()
    un = {k: [s[k] for s in samples] for k in keys}
    data = {k: v for k, v in un.items() if k in ('task_name', 'video_name', 'task_indices')}
    data['lengths'] = torch.LongTensor([f.size(0) for f in un['features']])
    for k in ('gt_single', 'features', 'constraints'):
        if k in un:
            data[k] = torch.nn.utils.rnn.pad_sequence(un[k], batch_first=True, padding_value=0)
    return data


class _Loader:
    """Same surface as the torch DataLoader the reference builds (models/model.py:66-77): `dataset`, `batch_sampler`
    (an iterable of lists of dataset keys), `collate_fn`."""

    def __init__(self, datasplit, batches):
        self.datasplit, self._batches = datasplit, batches
        self.dataset = datasplit
        self.batch_sampler = batches
        self.collate_fn = padding_colate

    def __len__(self):
        return len(self._batches)

    def __iter__(self):
        for ids in self._batches:
            yield padding_colate([self.datasplit[i] for i in ids])


def make_data_loader(args, datasplit, shuffle, batch_by_task, batch_size=1):
    return _Loader(datasplit, datasplit.batches(batch_size, batch_by_task, shuffle))


def make_crosstask_like(n_tasks=3, steps_per_task=(3, 5), n_videos=12, feature_dim=16, frames=(40, 80), max_seg=12,
                        sep=2.0, narration=False, seed=0, allow_short=False):
    """CrossTask-shaped split under --task_specific_steps --annotate_background_with_previous:
    task t owns the chain [bkg_0, step_1, bkg_1, ..., step_s, bkg_s] of consecutive global ids."""
    rng = np.random.RandomState(seed)
    chains, indices_by_task, background = {}, {}, []
    nxt = 0
    for t in range(n_tasks):
        s = int(rng.randint(steps_per_task[0], steps_per_task[1] + 1))
        ch = list(range(nxt, nxt + 2 * s + 1))
        nxt += 2 * s + 1
        chains[t] = ch
        indices_by_task[t] = ch
        background.extend(ch[0::2])
    n_classes = nxt
    mu = rng.randn(n_classes, feature_dim) * sep / np.sqrt(feature_dim)
    for t, ch in chains.items():  # backgrounds of a task look alike
        mu[ch[0::2]] = mu[ch[0]]
    videos = []
    for i in range(n_videos):
        t = i % n_tasks
        ch = chains[t]
        T = int(rng.randint(frames[0], frames[1] + 1))
        if not allow_short:
            T = max(T, len(ch))
        if T < len(ch):  # a video shorter than its task chain: one frame per class until it ends (semimarkov.py:135-147)
            lab = np.asarray(ch[:T])
        else:
            cuts = np.sort(rng.choice(np.arange(1, T), size=len(ch) - 1, replace=False))
            seg_len = np.diff(np.concatenate([[0], cuts, [T]]))
            lab = np.repeat(ch, seg_len)
        x = mu[lab] + rng.randn(T, feature_dim)
        v = dict(task_name=t, video_name='vid%04d' % i, features=torch.from_numpy(x).float(),
                 gt_single=torch.from_numpy(lab).long(), constraints=None)
        if narration:
            steps = ch[1::2]
            cons = np.zeros((T, len(steps)), dtype=np.float32)
            for j, c in enumerate(steps):
                where = np.flatnonzero(lab == c)
                if where.size == 0:
                    continue
                lo = max(0, where[0] - int(rng.randint(0, max_seg)))
                hi = min(T, where[-1] + 1 + int(rng.randint(0, max_seg)))
                cons[lo:hi, j] = 1
            v['constraints'] = torch.from_numpy(cons)
        videos.append(v)
    corpus = SyntheticCorpus(n_classes, indices_by_task, background, annotate_background_with_previous=True)
    return SyntheticDatasplit(corpus, videos, feature_dim, chains=chains)


def make_supervised_like(n_tasks=2, steps_per_task=(4, 6), n_videos=10, feature_dim=16, frames=(60, 120), sep=2.5, seed=0):
    """Split for the supervised S6 flow (README.md:43: no --task_specific_steps, one shared background class 0):
    task t owns a set of step classes; a video visits its task's steps in order with background in between, steps
    may repeat or be skipped."""
    rng = np.random.RandomState(seed)
    indices_by_task, nxt = {}, 1
    for t in range(n_tasks):
        s = int(rng.randint(steps_per_task[0], steps_per_task[1] + 1))
        indices_by_task[t] = [0] + list(range(nxt, nxt + s))
        nxt += s
    n_classes = nxt
    mu = rng.randn(n_classes, feature_dim) * sep / np.sqrt(feature_dim)
    videos = []
    for i in range(n_videos):
        t = i % n_tasks
        steps = indices_by_task[t][1:]
        T = int(rng.randint(frames[0], frames[1] + 1))
        seq = [0]
        for c in steps:
            if rng.rand() < 0.85:
                seq += [c, 0]
        cuts = np.sort(rng.choice(np.arange(1, T), size=len(seq) - 1, replace=False)) if len(seq) > 1 else np.array([], dtype=int)
        seg_len = np.diff(np.concatenate([[0], cuts, [T]]))
        lab = np.repeat(seq, seg_len)
        x = mu[lab] + rng.randn(T, feature_dim)
        videos.append(dict(task_name=t, video_name='sup%04d' % i, features=torch.from_numpy(x).float(),
                           gt_single=torch.from_numpy(lab).long(), constraints=None))
    corpus = SyntheticCorpus(n_classes, indices_by_task, [0])
    return SyntheticDatasplit(corpus, videos, feature_dim, chains=None)
