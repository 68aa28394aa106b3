"""TEST / BASELINE INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

CPU port of the reference's OWN way of computing the hot path, used as the timed CPU baseline
(`cpu_baseline.kind = "port"`, and `bench.py --impl reference`): /root/reference cannot travel to
the GPU box and its pytorch-struct dependency is not installable, so this module restates, with
torch CPU ops and the same algorithmic structure (and therefore the same cost):

  * emission scoring: a Python loop over classes, each a MultivariateNormal with a D x D scale_tril
    (semimarkov_modules.py:324-362);
  * dense potentials (B, T, K, C+1, C+1) with EOS augmentation and unfold-based sliding sums, one
    pass per k (semimarkov_modules.py:26-39, 416-523);
  * pytorch-struct's sequential DP, marginals / argmax by autograd (oracle/torch_struct_shim.py);
  * logZ.mean().backward() for the four parameter gradients (semimarkov.py:259-286) and
    argmax -> from_parts for Viterbi (semimarkov_modules.py:677-679).

Its numbers are checked against oracle/hsmm_oracle.py in tests/test_oracle_golden.py::test_reference_port_matches_oracle.
bench.py uses it only when the reference's own sources (oracle/_ref) are absent.
"""
import torch
import torch.nn.functional as F
from torch.distributions import MultivariateNormal

from .torch_struct_shim import SemiMarkovCRF

BIG_NEG = -1e9


def sliding_sum(x, k):
    b = x.size(0)
    if k == 1:
        return x
    win = F.unfold(x.unsqueeze(1), kernel_size=(k, 1), padding=(k, 0)).reshape(b, k, -1, x.size(-1))
    return win.sum(dim=1)[:, k:-1, :]


def emission_log_probs(features, means, cov, constraints=None):
    scale_tril = cov.sqrt()
    B = features.size(0)
    cols = []
    for c in range(means.size(0)):
        dist = MultivariateNormal(loc=means[c].unsqueeze(0).expand(B, -1), scale_tril=scale_tril)
        cols.append(dist.log_prob(features.transpose(0, 1)).transpose(0, 1).unsqueeze(-1))
    elp = torch.cat(cols, dim=2)
    return elp if constraints is None else elp + constraints


def dense_potentials(trans, em, init, len_scores, lengths, allowed_ends=None):
    b, n1, c1 = em.shape
    K = len_scores.size(0)
    if K > n1:
        K = n1
        len_scores = len_scores[:K]
    N, C = n1 + 1, c1 + 1
    tr = torch.full((b, C, C), BIG_NEG)
    tr[:, :c1, :c1] = trans
    if allowed_ends is None:
        tr[:, c1, :] = 0
    else:
        for i, ends in enumerate(allowed_ends):
            tr[i, c1, ends] = 0
    ini = torch.full((b, C), BIG_NEG)
    ini[:, :c1] = init
    ls = torch.full((b, K, C), BIG_NEG)
    ls[:, :, :c1] = len_scores
    ls[:, 1 if K > 1 else 0, c1] = 0
    ea = torch.full((b, N, C), BIG_NEG)
    for i, ln in enumerate(lengths):
        ea[i, :ln, :c1] = em[i, :ln]
        ea[i, ln, c1] = 0
    scores = torch.zeros(b, N - 1, K, C, C)
    scores += tr.view(b, 1, 1, C, C)
    scores[:, 0] += ini.view(b, 1, 1, C)
    scores += ls.view(b, 1, K, 1, C)
    for k in range(1, K):
        summed = sliding_sum(ea, k).view(b, N, 1, C)
        for i in range(b):
            ln = int(lengths[i]) + 1
            scores[i, :ln - 1, k] += summed[i, :ln - 1]
            scores[i, ln - 1 - k, k] += ea[i, ln - 1].view(C, 1)
    return scores


class ReferencePort:
    """Parameters in the reference's state_dict layout (full class set = the batch's valid classes)."""

    def __init__(self, means, cov_diag, trans_logits, init_logits, log_rates, max_k, trans_mask=None, init_mask=None):
        self.means = means.clone().requires_grad_(True)
        self.cov = torch.diag(cov_diag)
        self.trans_logits = trans_logits.clone().requires_grad_(True)
        self.init_logits = init_logits.clone().requires_grad_(True)
        self.log_rates = log_rates.clone().requires_grad_(True)
        self.max_k = max_k
        self.trans_mask, self.init_mask = trans_mask, init_mask

    def _scores(self, features, lengths, constraints=None, allowed_ends=None):
        il = self.init_logits if self.init_mask is None else self.init_logits.masked_fill(self.init_mask, BIG_NEG)
        tl = self.trans_logits if self.trans_mask is None else self.trans_logits.masked_fill(self.trans_mask, BIG_NEG)
        init, trans = F.log_softmax(il, dim=0), F.log_softmax(tl, dim=0)
        k = torch.arange(self.max_k, dtype=torch.float32).unsqueeze(-1)
        lenp = torch.distributions.Poisson(torch.exp(self.log_rates)).log_prob(k.expand(self.max_k, self.log_rates.numel()))
        elp = emission_log_probs(features, self.means, self.cov, constraints)
        return dense_potentials(trans, elp, init, lenp, lengths, allowed_ends)

    def train_step(self, features, lengths, constraints=None, allowed_ends=None):
        """logZ.mean() and its gradients, as loss.backward() produces them in the reference."""
        for p in (self.means, self.trans_logits, self.init_logits, self.log_rates):
            p.grad = None
        scores = self._scores(features, lengths, constraints, allowed_ends)
        ll = SemiMarkovCRF(scores, lengths=lengths + 1).partition.mean()
        ll.backward()
        return float(ll), dict(gaussian_means=self.means.grad, transition_logits=self.trans_logits.grad,
                               init_logits=self.init_logits.grad, poisson_log_rates=self.log_rates.grad)

    def viterbi(self, features, lengths, constraints=None, allowed_ends=None):
        scores = self._scores(features, lengths, constraints, allowed_ends)
        dist = SemiMarkovCRF(scores, lengths=lengths + 1)
        spans, _ = dist.struct.from_parts(dist.argmax)
        return spans
