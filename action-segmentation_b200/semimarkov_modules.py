"""Drop-in `SemiMarkovModule` for the reference's `--classifier semimarkov` path.

Same constructor, parameters (state_dict keys), flags and method signatures as
/root/reference/src/models/semimarkov/semimarkov_modules.py:52-696, but `log_likelihood` and `viterbi`
run on libhsmm_b200.so instead of building the (B,T,K,C+1,C+1) potentials (`log_hsmm`, :416-523) and
calling pytorch-struct.  What stays in torch is the part the reference also keeps as tiny tensor
ops: masking + log_softmax of the init/transition logits (:284-322), the Poisson length table
(:383-414) and index bookkeeping (valid classes, merged classes, allowed ends).  Those are
differentiated by autograd; the DP gradients come from the library's backward kernel.

Not supported (out of scope): the NICE feature projector
(`--sm_feature_projection`) and the component model.
"""
import pickle
from typing import Dict, Set

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import hsmm
from . import semimarkov_utils
from ._lib import HsmmError

BIG_NEG = -1e9  # semimarkov_modules.py:20


def all_equal(xs):
    xs = list(xs)
    return all(x == xs[0] for x in xs[1:])


class HsmmScores:
    """What `score_features` hands to the DP in place of the dense potential tensor."""

    def __init__(self, em, rowterm, offset, init, trans, lenp, end, lengths_i32, order, C, sparse=None):
        self.em, self.rowterm, self.offset = em, rowterm, offset
        self.init, self.trans, self.lenp, self.end = init, trans, lenp, end
        self.lengths_i32, self.order, self.C, self.sparse = lengths_i32, order, C, sparse

    @property
    def elp(self):
        """(B, T, C) emission log-probabilities as the reference's emission_log_probs returns them."""
        return self.em[:, :, :self.C] + self.rowterm.unsqueeze(-1)


class SemiMarkovModule(nn.Module):
    @classmethod
    def add_args(cls, parser):
        # semimarkov_modules.py:53-65 (NICETrans flags omitted: projector out of scope)
        parser.add_argument('--sm_max_span_length', type=int, default=20)
        parser.add_argument('--sm_supervised_state_smoothing', type=float, default=1e-2)
        parser.add_argument('--sm_supervised_length_smoothing', type=float, default=1e-1)
        parser.add_argument('--sm_supervised_method',
                            choices=['closed-form', 'gradient-based', 'closed-then-gradient'],
                            default='closed-form')
        parser.add_argument('--sm_feature_projection', action='store_true', help='use a flow (not supported)')
        parser.add_argument('--sm_init_non_projection_parameters_from')

    def __init__(self, args, n_classes, n_dims,
                 allow_self_transitions=False,
                 allowed_starts: Set[int] = None,
                 allowed_transitions: Dict[int, Set[int]] = None,
                 allowed_ends: Set[int] = None,
                 merge_classes: Dict[int, int] = None):
        super(SemiMarkovModule, self).__init__()
        self.args = args
        self.n_classes = n_classes
        self.input_feature_dim = n_dims
        self.feature_dim = n_dims
        self.allow_self_transitions = allow_self_transitions
        self.init_params()
        if allowed_starts is not None:
            assert allowed_transitions is not None
            self.set_transition_constraints(allowed_starts, allowed_transitions, allowed_ends)
        else:
            self.remove_transition_constraints()
        if getattr(args, 'sm_feature_projection', False):
            raise NotImplementedError("--sm_feature_projection (NICE flow) is outside the B200 hot path")
        self.feature_projector = None
        init_from = getattr(args, 'sm_init_non_projection_parameters_from', None)
        if init_from is not None:
            # semimarkov_modules.py:90-94: start from the parameters of a pickled SemiMarkovModel
            print("loading all non-flow parameters from {}".format(init_from))
            with open(init_from, 'rb') as f:
                sm = pickle.load(f)
            self.init_nonproject_parameters(sm.model)
        self.max_k = args.sm_max_span_length
        self._merge_classes = merge_classes
        self.kl = None
        self._sparse_cache = {}

    @property
    def merge_classes(self):
        return getattr(self, '_merge_classes', None)

    def init_nonproject_parameters(self, model):
        # semimarkov_modules.py:125-129 (there is no feature projector here, so nothing may be missing)
        assert isinstance(model, SemiMarkovModule)
        incompatible = self.load_state_dict(model.state_dict(), strict=False)
        assert not incompatible.unexpected_keys, incompatible.unexpected_keys
        assert not incompatible.missing_keys, incompatible.missing_keys

    def init_params(self):
        # semimarkov_modules.py:142-159
        self.poisson_log_rates = nn.Parameter(torch.zeros(self.n_classes).float(), requires_grad=True)
        self.gaussian_means = nn.Parameter(torch.zeros(self.n_classes, self.feature_dim).float(), requires_grad=True)
        # shared, tied, diagonal covariance matrix
        self.gaussian_cov = nn.Parameter(torch.eye(self.feature_dim).float(), requires_grad=False)
        # target x source
        self.transition_logits = nn.Parameter(torch.zeros(self.n_classes, self.n_classes).float(), requires_grad=True)
        self.init_logits = nn.Parameter(torch.zeros(self.n_classes).float(), requires_grad=True)
        torch.nn.init.uniform_(self.init_logits, 0, 1)

    def flatten_parameters(self):
        pass

    def remove_transition_constraints(self):
        self.__dict__['_sparse_cache'] = {}
        self.transition_constraints = None
        self.init_constraints = None
        self.allowed_ends = None

    def set_transition_constraints(self, allowed_starts, allowed_transitions, allowed_ends):
        # semimarkov_modules.py:169-193
        self.__dict__['_sparse_cache'] = {}
        init_c = torch.full((self.n_classes,), 1, dtype=torch.bool)
        assert all(x >= 0 for x in allowed_starts)
        init_c[torch.LongTensor(list(sorted(allowed_starts)))] = 0
        self.init_constraints = nn.Parameter(init_c, requires_grad=False)
        trans_c = torch.full((self.n_classes, self.n_classes), 1, dtype=torch.bool)
        for src, targets in allowed_transitions.items():
            for tgt in targets:
                trans_c[tgt, src] = 0
        self.transition_constraints = nn.Parameter(trans_c, requires_grad=False)
        self.allowed_ends = allowed_ends

    # ------------------------------------------------------------------------------------------
    # supervised closed form  (semimarkov_modules.py:195-256, semimarkov_utils.py:66-126)
    # ------------------------------------------------------------------------------------------
    def _gaussian_stats(self, feature_list, label_list):
        """Class means r^T X / n_c and tied diagonal variance E[x^2] - E[x]^2 + 1e-6 (biased), the
        moments sklearn's GaussianMixture._initialize produces for one-hot responsibilities."""
        dev = self.gaussian_means.device
        if dev.type != 'cuda':
            raise HsmmError("fit_supervised computes its feature statistics on the GPU: move the module to CUDA first")
        X = torch.cat([f.to(dev, torch.float32) for f in feature_list], dim=0).unsqueeze(0).contiguous()
        lab = torch.cat([l.to(dev) for l in label_list], dim=0).to(torch.int32).unsqueeze(0).contiguous()
        n = X.shape[1]
        lengths = torch.tensor([n], device=dev, dtype=torch.int32)
        onehot = hsmm.onehot_weights(lab, self.n_classes, lengths)
        wx, wsum = hsmm.weighted_feature_sums(X, onehot, self.n_classes, lengths)
        sx, sx2 = hsmm.feature_moments(X, lengths)
        eps = 10 * np.finfo(np.float64).eps
        means = wx.double() / (wsum.double() + eps)[:, None]
        nk = float(n) + eps
        mean_all = sx / nk
        var = sx2 / nk - 2 * mean_all * sx / nk + mean_all ** 2 + 1e-6
        return means, var

    def fit_supervised(self, feature_list, label_list):
        if self.feature_projector is not None:
            raise NotImplementedError("fit_supervised closed form with feature projector")
        if self.transition_constraints is not None or self.init_constraints is not None:
            raise NotImplementedError("fit_supervised closed form with constrained state transitions")
        a = self.args
        stats = semimarkov_utils.span_count_stats(label_list, self.n_classes, self.max_k)
        if self.merge_classes is not None:
            label_list_merged = [
                torch.as_tensor([self.merge_classes[int(ix)] for ix in labels], dtype=torch.long) for labels in label_list
            ]
            stats_merged = semimarkov_utils.span_count_stats(label_list_merged, self.n_classes, self.max_k)
        else:
            label_list_merged, stats_merged = label_list, stats
        means, var = self._gaussian_stats(feature_list, label_list_merged)

        # transition probs use unmerged classes
        init_probs = (stats['span_start_counts'] + a.sm_supervised_state_smoothing) / float(
            stats['instance_count'] + a.sm_supervised_state_smoothing * self.n_classes)
        init_probs[np.isnan(init_probs)] = 0
        self.init_logits.data.copy_(torch.from_numpy(init_probs).to(self.init_logits.device).log())
        smoothed = stats['span_transition_counts'] + a.sm_supervised_state_smoothing
        trans_probs = smoothed / smoothed.sum(axis=0)[None, :]
        trans_probs[np.isnan(trans_probs)] = 0
        self.transition_logits.data.copy_(torch.from_numpy(trans_probs).to(self.transition_logits.device).log())
        # lengths and emissions use merged classes
        mean_lengths = (stats_merged['span_lengths'] + a.sm_supervised_length_smoothing) / (
            stats_merged['span_counts'] + a.sm_supervised_length_smoothing)
        self.poisson_log_rates.data.copy_(torch.from_numpy(mean_lengths).to(self.poisson_log_rates.device).log())
        self.gaussian_means.data.copy_(means.float())
        self.gaussian_cov.data.copy_(torch.diag(var.float()))

    def initialize_gaussian_from_feature_list(self, features):
        # semimarkov_modules.py:263-274: every class mean = global mean, cov = diag(unbiased variance)
        dev = self.gaussian_means.device
        if dev.type != 'cuda':
            raise HsmmError("initialize_gaussian computes its feature statistics on the GPU: move the module to CUDA first")
        X = torch.cat([f.to(dev, torch.float32) for f in features], dim=0).unsqueeze(0).contiguous()
        n = X.shape[1]
        assert X.shape[2] == self.feature_dim
        lengths = torch.tensor([n], device=dev, dtype=torch.int32)
        sx, sx2 = hsmm.feature_moments(X, lengths)
        mean = sx / n
        var = (sx2 - n * mean * mean) / (n - 1)
        self.gaussian_means.data.copy_(mean.float().unsqueeze(0).expand(self.n_classes, self.feature_dim))
        self.gaussian_cov.data = torch.diag(var.float())

    def initialize_gaussian(self, data, lengths):
        batch_size = data.size(0)
        assert lengths.size(0) == batch_size
        self.initialize_gaussian_from_feature_list([data[i, :int(lengths[i])] for i in range(batch_size)])

    # ------------------------------------------------------------------------------------------
    # parameter -> score transforms (tiny; autograd-differentiated)
    # ------------------------------------------------------------------------------------------
    def initial_log_probs(self, valid_classes):
        logits = self.init_logits
        if self.init_constraints is not None:
            logits = logits.masked_fill(self.init_constraints, BIG_NEG)
        if valid_classes is not None:
            logits = logits[valid_classes]
        return F.log_softmax(logits, dim=0)

    def transition_log_probs(self, valid_classes):
        transition_logits = self.transition_logits
        if self.transition_constraints is not None:
            transition_logits = transition_logits.masked_fill(self.transition_constraints, BIG_NEG)
        if valid_classes is not None:
            transition_logits = transition_logits[valid_classes][:, valid_classes]
            n_classes = len(valid_classes)
        else:
            n_classes = self.n_classes
        if self.allow_self_transitions:
            masked = transition_logits
        else:
            masked = transition_logits.masked_fill(
                torch.eye(n_classes, device=self.transition_logits.device).bool(), BIG_NEG)
        # indexed [to_state, from_state]: each column is normalised
        return F.log_softmax(masked, dim=0)

    def _task_cache(self, valid_classes, device):
        """Index bookkeeping of one valid-class set (= one task, corpus.py:326-329), built once on the host and
        kept on the device: class indices (merged for emission/length parameters), the local->global id table
        for decoding, the sparse-transition hint.  No per-call host<->device synchronisation."""
        vc_host = None if valid_classes is None else [int(x) for x in valid_classes.detach().cpu().tolist()]
        key = (None if vc_host is None else tuple(vc_host), str(device))
        cache = self.__dict__.setdefault('_sparse_cache', {})
        if key not in cache:
            ids = list(range(self.n_classes)) if vc_host is None else vc_host
            merged = ids if self.merge_classes is None else [self.merge_classes[i] for i in ids]
            allowed = None
            if self.transition_constraints is not None:
                allowed = ~self.transition_constraints.detach().cpu()
                if vc_host is not None:
                    vc = torch.as_tensor(vc_host, dtype=torch.long)
                    allowed = allowed[vc][:, vc]
                if not self.allow_self_transitions:
                    allowed = allowed & ~torch.eye(allowed.shape[0], dtype=torch.bool)
            cache[key] = dict(
                ids_host=ids,
                idx=None if vc_host is None else torch.as_tensor(ids, dtype=torch.long).to(device),
                merged_idx=torch.as_tensor(merged, dtype=torch.long).to(device),
                decode_ids=torch.as_tensor(ids + [self.n_classes], dtype=torch.int32).to(device),
                sparse=None if allowed is None else hsmm.sparse_transition_lists(allowed, device))
        return cache[key]

    def _class_indices(self, valid_classes, device):
        return self._task_cache(valid_classes, device)['merged_idx']

    def _length_log_probs_with_rates(self, log_rates):
        # semimarkov_modules.py:383-398: Poisson(exp(log_rate)).log_prob(k), k = 0..max_k-1
        n_classes = log_rates.size(-1)
        max_length = self.max_k
        if max_length == 1:
            return torch.tensor([0.0, -1000.0], device=log_rates.device).unsqueeze(-1).expand(2, n_classes)
        k = torch.arange(max_length, device=log_rates.device, dtype=log_rates.dtype).unsqueeze(-1)
        return k * log_rates.unsqueeze(0) - torch.exp(log_rates).unsqueeze(0) - torch.lgamma(k + 1)

    def length_log_probs(self, valid_classes):
        idx = self._class_indices(valid_classes, self.poisson_log_rates.device)
        return self._length_log_probs_with_rates(self.poisson_log_rates[idx])

    def emission_log_probs(self, features, valid_classes, constraints):
        """(B, T, C) log N(x; mu_c, diag(cov)) (+ constraints) -- semimarkov_modules.py:364-381,
        computed by hsmm_emission.  No gradient flows through this debugging view."""
        idx = self._class_indices(valid_classes, self.gaussian_means.device)
        B, T, _ = features.shape
        lengths_i32 = torch.full((B,), T, device=features.device, dtype=torch.int32)
        em, rowterm, _ = hsmm.emission_scores(features, self.gaussian_means[idx], torch.diagonal(self.gaussian_cov),
                                              constraints, lengths_i32)
        return em[:, :, :len(idx)] + rowterm.unsqueeze(-1)

    # ------------------------------------------------------------------------------------------
    def add_eos(self, spans, lengths):
        b, N = spans.size()
        augmented = torch.cat([spans, torch.full([b, 1], -1, device=spans.device, dtype=torch.long)], dim=1)
        augmented[torch.arange(b), lengths] = self.n_classes
        return augmented

    def trim(self, spans, lengths, check_eos=False):
        # lengths should be the lengths NOT including any eos symbol at the end
        return [spans[i, :int(lengths[i])] for i in range(spans.size(0))]

    @property
    def batched_scores(self):
        return False

    def set_z(self, features, lengths, use_mean=False):
        self.kl = torch.zeros(features.size(0), device=features.device, requires_grad=True)

    def _end_scores(self, ids_host, batch_size, additional_allowed_ends_per_instance, device):
        """EOS row of log_hsmm's augmented transitions (semimarkov_modules.py:462-471, 566-577); built in pinned
        host memory and copied asynchronously."""
        if self.allowed_ends is None:
            return None
        base = torch.tensor([0.0 if ix in self.allowed_ends else BIG_NEG for ix in ids_host])
        end = base.unsqueeze(0).repeat(batch_size, 1)
        if additional_allowed_ends_per_instance is not None:
            pos = {ix: i for i, ix in enumerate(ids_host)}
            for b, extra in enumerate(additional_allowed_ends_per_instance):
                for x in extra:
                    if int(x) in pos:
                        end[b, pos[int(x)]] = 0.0
        assert bool((end == 0).any(dim=1).all()), "no allowed end state among the valid classes"
        return end.pin_memory().to(device, non_blocking=True)

    def __getstate__(self):
        d = super().__getstate__() if hasattr(super(), '__getstate__') else dict(self.__dict__)
        d = dict(d)
        d['_sparse_cache'] = {}
        return d

    def _scores(self, features, lengths, valid_classes, additional_allowed_ends_per_instance):
        dev = features.device
        if dev.type != 'cuda':
            raise HsmmError("the HSMM path runs on CUDA only (no CPU fallback); call .cuda() on the model and inputs")
        tc = self._task_cache(valid_classes, dev)
        idx = tc['merged_idx']
        C = len(tc['ids_host'])
        T = features.size(1)
        lenp = self._length_log_probs_with_rates(self.poisson_log_rates[idx])
        K = lenp.size(0)
        if K > T:  # semimarkov_modules.py:450-452
            K = T
            lenp = lenp[:K]
        if K < 2:
            raise HsmmError("padded batch length %d leaves no usable segment length" % T)
        lengths_i32, order = hsmm.prepare_lengths(lengths, dev)
        end = self._end_scores(tc['ids_host'], features.size(0), additional_allowed_ends_per_instance, dev)
        return dict(means=self.gaussian_means[idx], cov_diag=torch.diagonal(self.gaussian_cov),
                    init=self.initial_log_probs(tc['idx']), trans=self.transition_log_probs(tc['idx']),
                    lenp=lenp, end=end, lengths_i32=lengths_i32, order=order, C=C, valid_classes=tc['idx'],
                    sparse=tc['sparse'], decode_ids=tc['decode_ids'])

    def score_features(self, features, lengths, valid_classes, add_eos, use_mean_z,
                       additional_allowed_ends_per_instance=None, constraints=None, return_elp=False):
        """Returns (HsmmScores, log_det[, elp]).  The reference returns the dense potential tensor
        here (semimarkov_modules.py:553-595); this implementation never builds it."""
        assert add_eos, "only the add_eos=True formulation (the one the reference uses) is implemented"
        self.set_z(features, lengths, use_mean=use_mean_z)
        s = self._scores(features, lengths, valid_classes, additional_allowed_ends_per_instance)
        with torch.no_grad():
            em, rowterm, offset = hsmm.emission_scores(features, s['means'], s['cov_diag'], constraints, s['lengths_i32'])
        scores = HsmmScores(em, rowterm, offset, s['init'].detach(), s['trans'].detach(), s['lenp'].detach(), s['end'],
                            s['lengths_i32'], s['order'], s['C'], s['sparse'])
        scores.decode_ids = s['decode_ids']
        log_det = torch.zeros(features.size(0), device=features.device, requires_grad=False)
        if return_elp:
            return scores, log_det, scores.elp
        return scores, log_det

    def _valid_classes(self, valid_classes_per_instance):
        if valid_classes_per_instance is None:
            return None, self.n_classes
        assert all_equal(set(vc.detach().cpu().numpy()) for vc in valid_classes_per_instance), \
            "must have same valid_classes for all instances in the batch"
        vc = valid_classes_per_instance[0]
        return vc, len(vc)

    def log_likelihood(self, features, lengths, valid_classes_per_instance, spans=None, add_eos=True, use_mean_z=False,
                       additional_allowed_ends_per_instance=None, constraints=None):
        """semimarkov_modules.py:597-658: mean over the batch of logZ (spans=None), of the gold-path
        score (generative) or of score - logZ (--sm_train_discriminatively)."""
        assert add_eos, "only add_eos=True is implemented"
        valid_classes, C = self._valid_classes(valid_classes_per_instance)
        self.set_z(features, lengths, use_mean=use_mean_z)
        s = self._scores(features, lengths, valid_classes, additional_allowed_ends_per_instance)
        args = (features, s['means'], s['cov_diag'], constraints, s['init'], s['trans'], s['lenp'], s['end'],
                s['lengths_i32'])
        log_det = torch.zeros(features.size(0), device=features.device, requires_grad=False)
        if spans is None:
            logz, _, _ = hsmm.HsmmLogZ.apply(*args, s['order'], s['sparse'])
            return logz.mean(), log_det.mean()
        # gold spans arrive in global class ids; map to positions in valid_classes
        dev = features.device
        spans = spans.to(dev)
        lut = torch.full((self.n_classes + 1,), -2, dtype=torch.long, device=dev)  # -2: not a valid class -> NaN score
        vc = torch.arange(self.n_classes, device=dev) if s['valid_classes'] is None else s['valid_classes']
        lut[vc] = torch.arange(len(vc), device=dev)
        local = torch.where(spans >= 0, lut[spans.clamp(min=0)], torch.full_like(spans, -1)).to(torch.int32).contiguous()
        score = hsmm.HsmmGoldScore.apply(*args, local)
        if getattr(self.args, 'sm_train_discriminatively', False):
            logz, _, _ = hsmm.HsmmLogZ.apply(*args, s['order'], s['sparse'])
            return (score - logz).mean(), log_det.mean()
        return score.mean(), log_det.mean()

    # ------------------------------------------------------------------------------------------
    # closed-form EM (optional trainer beside the reference's Adam-on-logZ)
    # ------------------------------------------------------------------------------------------
    STAT_KEYS = ('wx', 'wsum', 'init', 'trans', 'len_num', 'len_den', 'logz', 'n')

    def expected_statistics(self, features, lengths, valid_classes_per_instance,
                            additional_allowed_ends_per_instance=None, constraints=None):
        """E-step of one batch with the same kernels as `log_likelihood().backward()`: the expected counts the
        reference only ever sees as gradients (semimarkov.py:284-286), in GLOBAL class layout so that batches of
        different tasks (and ranks: `distributed.allreduce_stats(pack_statistics(...))`) add up.
          wx (n_classes, D)   sum_t post[t,c] x_t      wsum (n_classes)  sum_t post[t,c]      (merged classes)
          init (n_classes)    E[first segment is c]    trans (n_classes, n_classes) [to, from] expected transitions
          len_num / len_den   sum_k k E_len[k,c] / sum_k E_len[k,c]                            (merged classes)
          logz  sum of logZ_b   n  number of videos"""
        valid_classes, C = self._valid_classes(valid_classes_per_instance)
        dev = features.device
        with torch.no_grad():
            s = self._scores(features, lengths, valid_classes, additional_allowed_ends_per_instance)
            em, rowterm, offset = hsmm.emission_scores(features, s['means'], s['cov_diag'], constraints, s['lengths_i32'])
            xp = constraints is not None
            pred, succ = (None, None) if s['sparse'] is None else s['sparse']
            logz, saved = hsmm.logz_forward(em, C, s['init'], s['trans'], s['lenp'], s['end'], offset, s['lengths_i32'],
                                            s['order'], trans_pred=pred, f64_state=xp)
            g = torch.ones(features.size(0), device=dev)
            d_init, d_trans, d_len, d_em = hsmm.logz_backward(em, C, s['init'], s['trans'], s['lenp'], s['end'],
                                                             s['lengths_i32'], s['order'], g, saved, trans_succ=succ,
                                                             f64_state=xp)
            wx, wsum = hsmm.weighted_feature_sums(features, d_em, C, s['lengths_i32'])
            tc = self._task_cache(valid_classes, dev)
            ids = torch.arange(self.n_classes, device=dev) if tc['idx'] is None else tc['idx']
            midx = tc['merged_idx']
            n, D = self.n_classes, self.feature_dim
            k = torch.arange(d_len.size(0), device=dev, dtype=torch.float32).unsqueeze(-1)
            stats = dict(wx=torch.zeros(n, D, device=dev).index_add_(0, midx, wx),
                         wsum=torch.zeros(n, device=dev).index_add_(0, midx, wsum),
                         init=torch.zeros(n, device=dev).index_add_(0, ids, d_init),
                         trans=torch.zeros(n, n, device=dev),
                         len_num=torch.zeros(n, device=dev).index_add_(0, midx, (d_len * k).sum(0)),
                         len_den=torch.zeros(n, device=dev).index_add_(0, midx, d_len.sum(0)),
                         logz=logz.sum().to(torch.float32).reshape(1),
                         n=torch.full((1,), float(features.size(0)), device=dev))
            stats['trans'][ids.unsqueeze(1), ids.unsqueeze(0)] += d_trans
        return stats

    @classmethod
    def pack_statistics(cls, stats):
        return torch.cat([stats[k].reshape(-1) for k in cls.STAT_KEYS])

    def statistics_size(self):
        n, D = self.n_classes, self.feature_dim
        return n * D + n + n + n * n + n + n + 2

    def unpack_statistics(self, buf):
        n, D = self.n_classes, self.feature_dim
        shapes = dict(wx=(n, D), wsum=(n,), init=(n,), trans=(n, n), len_num=(n,), len_den=(n,), logz=(1,), n=(1,))
        out, off = {}, 0
        for k in self.STAT_KEYS:
            m = int(np.prod(shapes[k]))
            out[k] = buf[off:off + m].view(*shapes[k])
            off += m
        return out

    @staticmethod
    def add_statistics(a, b):
        return b if a is None else {k: a[k] + b[k] for k in a}

    def em_update(self, stats, state_smoothing=0.0, length_smoothing=0.0, min_count=1e-6):
        """Closed-form M-step from (summed) `expected_statistics`: class means = wx / wsum (tied diagonal covariance
        stays fixed, as in the reference where `gaussian_cov` has requires_grad=False, semimarkov_modules.py:150-151),
        Poisson rate = expected mean segment length, transition columns and initial distribution = normalised
        expected counts over the unmasked entries.  Classes that received no mass keep their parameters.
        Emission/length parameters live at the merged class ids."""
        with torch.no_grad():
            wsum = stats['wsum']
            seen = wsum > min_count
            means = stats['wx'] / wsum.clamp(min=min_count).unsqueeze(-1)
            self.gaussian_means.data = torch.where(seen.unsqueeze(-1), means, self.gaussian_means.data)
            lden = stats['len_den']
            rate = (stats['len_num'] + length_smoothing) / (lden + length_smoothing).clamp(min=min_count)
            self.poisson_log_rates.data = torch.where(lden > min_count, rate.clamp(min=1e-3).log(), self.poisson_log_rates.data)
            tr = stats['trans'] + state_smoothing
            if self.transition_constraints is not None:
                tr = tr.masked_fill(self.transition_constraints, 0.0)
            if not self.allow_self_transitions:
                tr = tr.masked_fill(torch.eye(self.n_classes, device=tr.device).bool(), 0.0)
            col = tr.sum(dim=0, keepdim=True)
            new_t = torch.where(tr > 0, (tr / col.clamp(min=min_count)).clamp(min=1e-30).log(), torch.full_like(tr, -30.0))
            self.transition_logits.data = torch.where(col > min_count, new_t, self.transition_logits.data)
            ini = stats['init'] + state_smoothing
            if self.init_constraints is not None:
                ini = ini.masked_fill(self.init_constraints, 0.0)
            tot = ini.sum()
            if float(tot) > min_count:
                self.init_logits.data = torch.where(ini > 0, (ini / tot).clamp(min=1e-30).log(), torch.full_like(ini, -30.0))
        return float(stats['logz']) / max(1.0, float(stats['n']))

    @staticmethod
    def _to_host(t, non_blocking):
        if t is None:
            return None
        if not non_blocking:
            return t.cpu()
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        return h

    def viterbi(self, features, lengths, valid_classes_per_instance, add_eos=True, use_mean_z=False,
                additional_allowed_ends_per_instance=None, constraints=None, predict_single=False, return_elp=False,
                return_labels=False, non_blocking=False, return_spans=True):
        """semimarkov_modules.py:660-696: span-encoded predictions (b x T+1, CPU int64, global class
        ids, EOS = n_classes at position lengths[b]).  `return_labels=True` additionally returns the
        per-frame labels the kernel emits (replacing semimarkov_utils.spans_to_labels).
        `non_blocking=True` copies the results into pinned host memory asynchronously on the current
        stream: the caller synchronises the stream (or device) before reading them.
        `return_spans=False` skips the device->host copy of the span encoding (first result is None)."""
        assert add_eos, "only add_eos=True is implemented"
        valid_classes, C = self._valid_classes(valid_classes_per_instance)
        with torch.no_grad():
            scores, _ = self.score_features(features, lengths, valid_classes, add_eos=add_eos, use_mean_z=use_mean_z,
                                            additional_allowed_ends_per_instance=additional_allowed_ends_per_instance,
                                            constraints=constraints)
            spans, labels, _ = hsmm.viterbi_decode(scores.em, C, scores.init, scores.trans, scores.lenp, scores.end,
                                                   scores.offset, scores.lengths_i32, scores.order, scores.decode_ids,
                                                   want_labels=return_labels, want_score=False,
                                                   trans_pred=None if scores.sparse is None else scores.sparse[0])
        out = [self._to_host(spans, non_blocking) if return_spans else None]
        if return_elp:
            out.append(scores.elp)
        if return_labels:
            out.append(self._to_host(labels, non_blocking))
        return out[0] if len(out) == 1 else tuple(out)

    GROUP_MAX = 32  # hsmm_dp_grouped: batches per call

    def viterbi_batches(self, batches, return_labels=True, return_spans=True, non_blocking=True):
        """`viterbi` for a LIST of batches (dicts with features, lengths, valid_classes_per_instance and optionally
        additional_allowed_ends_per_instance, constraints): emission scoring per batch, then ONE grouped Viterbi launch
        per kernel family for up to 32 batches at a time (hsmm_dp_grouped) when the batches are inside its envelope --
        ordering-constrained transitions, max span <= 21, <= 32 valid classes, which is the reference's U7 setting --
        and per-batch launches otherwise.  This is the decode of a whole split (models/semimarkov/semimarkov.py:318-410
        loops over mini-batches of 5 videos).  Returns a list of (spans, labels) like `viterbi(return_labels=True)`."""
        prepared = []
        with torch.no_grad():
            for bt in batches:
                valid_classes, C = self._valid_classes(bt['valid_classes_per_instance'])
                scores, _ = self.score_features(bt['features'], bt['lengths'], valid_classes, add_eos=True, use_mean_z=True,
                                                additional_allowed_ends_per_instance=bt.get('additional_allowed_ends_per_instance'),
                                                constraints=bt.get('constraints'))
                prepared.append((scores, C))
            results = [None] * len(prepared)
            todo = [i for i, (sc, C) in enumerate(prepared)
                    if sc.sparse is not None and sc.lenp.shape[0] - 1 <= 20 and C <= 32]
            for a in range(0, len(todo), self.GROUP_MAX):
                chunk = todo[a:a + self.GROUP_MAX]
                res = hsmm.grouped_dp(0, [dict(em=prepared[i][0].em, C=prepared[i][1], init=prepared[i][0].init,
                                               trans=prepared[i][0].trans, lenp=prepared[i][0].lenp, end=prepared[i][0].end,
                                               offset=prepared[i][0].offset, lengths_i32=prepared[i][0].lengths_i32,
                                               order=prepared[i][0].order, trans_list=prepared[i][0].sparse[0],
                                               class_ids=prepared[i][0].decode_ids, want_labels=return_labels) for i in chunk])
                for i, (spans, labels, _) in zip(chunk, res):
                    results[i] = (spans, labels)
            for i, (sc, C) in enumerate(prepared):
                if results[i] is None:
                    spans, labels, _ = hsmm.viterbi_decode(sc.em, C, sc.init, sc.trans, sc.lenp, sc.end, sc.offset, sc.lengths_i32,
                                                           sc.order, sc.decode_ids, want_labels=return_labels, want_score=False,
                                                           trans_pred=None if sc.sparse is None else sc.sparse[0])
                    results[i] = (spans, labels)
        return [(self._to_host(sp, non_blocking) if return_spans else None,
                 self._to_host(lab, non_blocking) if return_labels else None) for sp, lab in results]

    def log_likelihood_and_viterbi(self, features, lengths, valid_classes_per_instance, add_eos=True,
                                   additional_allowed_ends_per_instance=None, constraints=None, non_blocking=False):
        """`log_likelihood(spans=None)` and `viterbi(return_labels=True)` of the same batch with ONE emission pass
        (the unsupervised trainer's step followed by the decode of the same videos, main.py:207-218).
        Returns (ll_mean, log_det_mean, spans, labels); ll_mean carries the same autograd graph as log_likelihood's."""
        assert add_eos, "only add_eos=True is implemented"
        valid_classes, C = self._valid_classes(valid_classes_per_instance)
        self.set_z(features, lengths, use_mean=False)
        s = self._scores(features, lengths, valid_classes, additional_allowed_ends_per_instance)
        with torch.no_grad():
            em_pack = hsmm.emission_scores(features, s['means'], s['cov_diag'], constraints, s['lengths_i32'])
        logz, _, _ = hsmm.HsmmLogZ.apply(features, s['means'], s['cov_diag'], constraints, s['init'], s['trans'], s['lenp'],
                                         s['end'], s['lengths_i32'], s['order'], s['sparse'], em_pack)
        with torch.no_grad():
            spans, labels, _ = hsmm.viterbi_decode(em_pack[0], C, s['init'], s['trans'], s['lenp'], s['end'], em_pack[2],
                                                   s['lengths_i32'], s['order'], s['decode_ids'], want_labels=True,
                                                   want_score=False, trans_pred=None if s['sparse'] is None else s['sparse'][0])
        log_det = torch.zeros(features.size(0), device=features.device, requires_grad=False)
        return logz.mean(), log_det.mean(), self._to_host(spans, non_blocking), self._to_host(labels, non_blocking)
