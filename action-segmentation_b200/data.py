"""Synthetic stand-in for the reference's data layer, exposing exactly the surface the HSMM wrapper
touches (src/data/corpus.py: Datasplit.__getitem__ :315-380, BatchSampler :613-644; src/data/crosstask.py:
get_allowed_starts_and_transitions :328-388; src/models/model.py: padding_colate :42-63,
make_data_loader :66-77).  The reference's datasets are not available offline; benchmarks and tests
use CrossTask-/Breakfast-shaped synthetic videos instead."""
import numpy as np
import torch


class SyntheticCorpus:
    def __init__(self, n_classes, indices_by_task, background_indices):
        self.n_classes = n_classes
        self._indices_by_task = indices_by_task
        self._background_indices = list(background_indices)

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

    def batches(self, batch_size, batch_by_task, shuffle, seed=0):
        idx = list(range(len(self.videos)))
        rng = np.random.RandomState(seed)
        if shuffle:
            rng.shuffle(idx)
        if not batch_by_task:
            return [idx[i:i + batch_size] for i in range(0, len(idx), batch_size)]
        by_task = {}
        for i in idx:
            by_task.setdefault(self.videos[i]['task_name'], []).append(i)
        out = []
        for t in sorted(by_task):
            ids = by_task[t]
            out.extend(ids[i:i + batch_size] for i in range(0, len(ids), batch_size))
        if shuffle:
            rng.shuffle(out)
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
    def __init__(self, datasplit, batches):
        self.datasplit, self._batches = datasplit, batches
        self.dataset = datasplit

    def __len__(self):
        return len(self._batches)

    def __iter__(self):
        for ids in self._batches:
            yield padding_colate([self.datasplit[i] for i in ids])


def make_data_loader(args, datasplit, shuffle, batch_by_task, batch_size=1):
    seed = getattr(args, 'seed', 0)
    return _Loader(datasplit, datasplit.batches(batch_size, batch_by_task, shuffle, seed))


def make_crosstask_like(n_tasks=3, steps_per_task=(3, 5), n_videos=12, feature_dim=16, frames=(40, 80), max_seg=12,
                        sep=2.0, narration=False, seed=0):
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
        T = max(T, len(ch))
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
                lo = max(0, where[0] - int(rng.randint(0, max_seg)))
                hi = min(T, where[-1] + 1 + int(rng.randint(0, max_seg)))
                cons[lo:hi, j] = 1
            v['constraints'] = torch.from_numpy(cons)
        videos.append(v)
    corpus = SyntheticCorpus(n_classes, indices_by_task, background)
    return SyntheticDatasplit(corpus, videos, feature_dim, chains=chains)
