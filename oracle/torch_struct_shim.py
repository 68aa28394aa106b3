"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

CPU restatement of the slice of pytorch-struct that the reference's HSMM path calls.

The reference pins ``harvardnlp/pytorch-struct@1c9b038a1bbece32fe8d2d46d9e3d7c09f4c08e7``
(/root/reference/README.md:21-22, env.yml:46).  That package is NOT vendored under
/root/reference and cannot be installed here (no network), so this file restates the
published algorithm of its ``SemiMarkov`` struct (the sequential, low-memory DP of that
era) and the ``SemiMarkovCRF`` distribution wrapper, anchored on the reference's call
sites:

  * semimarkov_modules.py:11     ``from torch_struct import SemiMarkovCRF``
  * semimarkov_modules.py:624    ``SemiMarkovCRF(scores, lengths=eos_lengths)``
  * semimarkov_modules.py:641    ``SemiMarkovCRF.struct.to_parts(spans, (C, K), lengths=...)``
  * semimarkov_modules.py:646    ``dist.log_prob(parts)``
  * semimarkov_modules.py:650-655 ``dist.event_shape``, ``dist.struct().score(...)``, ``dist.log_potentials``
  * semimarkov_modules.py:657    ``dist.partition``
  * semimarkov_modules.py:677-679 ``dist.argmax``, ``dist.struct.from_parts``
  * test_semimarkov.py:7,14,312-314 ``SemiMarkov(MaxSemiring).marginals / .from_parts``

PARITY STATUS: the DP arithmetic here is "parity unpinned" against the real
pytorch-struct binary (absent).  It is pinned instead by (i) the reference's own
known-answer test body (test_semimarkov.py:266-323, reproduced in
tests/golden/make_golden.py::case_known_answer -> tests/test_oracle_golden.py::test_known_answer), and (ii) brute-force enumeration of every segmentation
(oracle/hsmm_oracle.py::brute_force) for logZ, max score and marginals.

Semantics restated (edge[b, n, k, c2, c1]: a segment labelled c1 starts at n, has length
k, and the next segment, labelled c2, starts at n + k):

    beta[0][c]      = one
    alpha[n-1][k,c2]= (+)_{c1} beta[n-1][c1] (x) edge[n-1, k, c2, c1]
    beta[n][c2]     = (+)_{k=1..min(K-1, n)} alpha[n-k][k, c2]
    v[b]            = (+)_c beta[lengths[b]-1][c]

marginals / argmax are the autograd gradient of v.sum() w.r.t. edge.
"""
import torch


class LogSemiring:
    @staticmethod
    def one():
        return 0.0

    @staticmethod
    def sum(x, dim=-1):
        return torch.logsumexp(x, dim=dim)

    @staticmethod
    def dot(a, b):
        return torch.logsumexp(a + b, dim=-1)

    @staticmethod
    def prod(x, dim=-1):
        return x.sum(dim=dim)


class MaxSemiring:
    @staticmethod
    def one():
        return 0.0

    @staticmethod
    def sum(x, dim=-1):
        return torch.max(x, dim=dim)[0]

    @staticmethod
    def dot(a, b):
        return torch.max(a + b, dim=-1)[0]

    @staticmethod
    def prod(x, dim=-1):
        return x.sum(dim=dim)


class SemiMarkov:
    """Semi-Markov struct over edge potentials ``b x (N-1) x K x C x C``."""

    def __init__(self, semiring=LogSemiring):
        self.semiring = semiring

    # -- DP ---------------------------------------------------------------------------
    def _dp(self, edge, lengths=None):
        sr = self.semiring
        batch, n_1, K, C, C2 = edge.shape
        assert C == C2, "Transition shape doesn't match"
        N = n_1 + 1
        if lengths is None:
            lengths = torch.full((batch,), N, dtype=torch.long)
        lengths = [int(l) for l in lengths]
        assert max(lengths) <= N, "Length longer than edge scores"
        assert max(lengths) == N, "At least one in batch must be length N"

        beta = [edge.new_full((batch, C), sr.one())]
        alpha = []  # alpha[m]: (batch, K, C) -- segments starting at m
        for n in range(1, N):
            alpha.append(sr.dot(beta[n - 1].view(batch, 1, 1, C), edge[:, n - 1]))
            ks = range(1, min(K - 1, n) + 1)
            stacked = torch.stack([alpha[n - k][:, k] for k in ks], dim=-1)
            beta.append(sr.sum(stacked, dim=-1))
        final = torch.stack([beta[l - 1][i] for i, l in enumerate(lengths)], dim=0)
        return sr.sum(final, dim=-1), beta

    def sum(self, edge, lengths=None):
        return self._dp(edge, lengths)[0]

    def marginals(self, edge, lengths=None):
        with torch.enable_grad():
            if not edge.requires_grad:
                edge = edge.detach().requires_grad_(True)
            v, _ = self._dp(edge, lengths)
            (marg,) = torch.autograd.grad(v.sum(), edge, create_graph=False)
        return marg

    # -- conversions ------------------------------------------------------------------
    @staticmethod
    def to_parts(sequence, extra, lengths=None):
        """b x N span encoding (-1 = continuation) -> one-hot b x (N-1) x K x C x C."""
        C, K = extra
        batch, N = sequence.shape
        parts = torch.zeros(batch, N - 1, K, C, C, dtype=torch.long)
        for b in range(batch):
            last, c = None, None
            for n in range(N):
                sym = int(sequence[b, n])
                if sym == -1:
                    assert n != 0
                    continue
                if n != 0:
                    parts[b, last, n - last, sym, c] = 1
                last, c = n, sym
        return parts

    @staticmethod
    def from_parts(edge):
        """one-hot edges -> (b x N span encoding, (C, K))."""
        batch, n_1, K, C, _ = edge.shape
        labels = torch.full((batch, n_1 + 1), -1, dtype=torch.long)
        for b, n, k, c2, c1 in edge.nonzero().tolist():
            if n == 0:
                labels[b, 0] = c1
            labels[b, n + k] = c2
        return labels, (C, K)

    def score(self, potentials, parts, batch_dims=(0,)):
        s = potentials * parts
        batch = tuple(s.shape[b] for b in batch_dims)
        return s.reshape(batch + (-1,)).sum(dim=-1)


class SemiMarkovCRF:
    """Distribution wrapper with the attributes the reference reads."""

    struct = SemiMarkov

    def __init__(self, log_potentials, lengths=None):
        self.log_potentials = log_potentials
        self.lengths = lengths
        self.event_shape = log_potentials.shape[1:]
        self.batch_shape = log_potentials.shape[:1]

    @property
    def partition(self):
        return SemiMarkov(LogSemiring).sum(self.log_potentials, self.lengths)

    @property
    def argmax(self):
        return SemiMarkov(MaxSemiring).marginals(self.log_potentials, self.lengths)

    @property
    def marginals(self):
        return SemiMarkov(LogSemiring).marginals(self.log_potentials, self.lengths)

    def log_prob(self, value):
        d = value.dim()
        batch_dims = range(d - len(self.event_shape))
        v = SemiMarkov().score(self.log_potentials, value.type_as(self.log_potentials), batch_dims=batch_dims)
        return v - self.partition
