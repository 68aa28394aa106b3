"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

fp64 module-level oracle: the reference's ``SemiMarkovModule.log_likelihood`` / ``.viterbi``
(/root/reference/src/models/semimarkov/semimarkov_modules.py:597-696) restated on top of
oracle/hsmm_oracle.py.  Parameter gradients are obtained the way the product does it: the DP
yields expected counts (= d logZ / d scores) and the tiny parameter->score transforms
(log_softmax, Poisson log-pmf, Gaussian log-density) are differentiated by torch autograd in fp64.
"""
import numpy as np
import torch

from . import hsmm_oracle as O


class ModuleOracle:
    def __init__(self, params, max_k, init_constraints=None, transition_constraints=None,
                 allowed_ends=None, merge_classes=None, allow_self_transitions=True):
        """params: dict of numpy arrays with the reference's state_dict keys."""
        self.p = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in params.items()
                  if k in ("gaussian_means", "gaussian_cov", "transition_logits", "init_logits", "poisson_log_rates")}
        self.n_classes = self.p["init_logits"].numel()
        self.max_k = int(max_k)
        self.init_constraints = None if init_constraints is None else torch.tensor(np.asarray(init_constraints), dtype=torch.bool)
        self.transition_constraints = None if transition_constraints is None else torch.tensor(np.asarray(transition_constraints), dtype=torch.bool)
        self.allowed_ends = None if allowed_ends is None else set(int(x) for x in allowed_ends)
        self.merge = merge_classes
        self.allow_self = allow_self_transitions

    # -- scores with autograd ---------------------------------------------------------------
    def _scores(self, features, valid, constraints):
        p = {k: v.clone().requires_grad_(k != "gaussian_cov") for k, v in self.p.items()}
        vc = torch.arange(self.n_classes) if valid is None else torch.as_tensor(np.asarray(valid)).long()
        mc = vc if self.merge is None else torch.tensor([self.merge[int(i)] for i in vc])
        il = p["init_logits"]
        if self.init_constraints is not None:
            il = il.masked_fill(self.init_constraints, O.BIG_NEG)
        init = torch.log_softmax(il[vc], dim=0)
        tl = p["transition_logits"]
        if self.transition_constraints is not None:
            tl = tl.masked_fill(self.transition_constraints, O.BIG_NEG)
        tl = tl[vc][:, vc]
        if not self.allow_self:
            tl = tl.masked_fill(torch.eye(len(vc)).bool(), O.BIG_NEG)
        trans = torch.log_softmax(tl, dim=0)
        lr = p["poisson_log_rates"][mc]
        k = torch.arange(self.max_k, dtype=torch.float64)[:, None]
        lenp = k * lr[None] - torch.exp(lr)[None] - torch.lgamma(k + 1)
        x = torch.tensor(np.asarray(features), dtype=torch.float64)
        var = torch.diagonal(p["gaussian_cov"])
        mu = p["gaussian_means"][mc]
        diff = x[..., None, :] - mu
        em = -0.5 * (diff * diff / var).sum(-1) - 0.5 * torch.log(var).sum() - 0.5 * x.shape[-1] * O.LOG_2PI
        if constraints is not None:
            em = em + torch.tensor(np.asarray(constraints), dtype=torch.float64)
        return p, vc, init, trans, lenp, em

    def _ends(self, B, vc, addl_ends):
        if self.allowed_ends is None:
            return None
        ends = np.full((B, len(vc)), O.BIG_NEG)
        for b in range(B):
            extra = set() if addl_ends is None else set(int(x) for x in addl_ends[b])
            for i, c in enumerate(vc.tolist()):
                if c in (self.allowed_ends | extra):
                    ends[b, i] = 0.0
        return ends

    # -- public ----------------------------------------------------------------------------
    def log_likelihood(self, features, lengths, valid=None, addl_ends=None, constraints=None):
        """Returns dict(ll, logz (B,), grads {state_dict key: array}, elp, counts)."""
        p, vc, init, trans, lenp, em = self._scores(features, valid, constraints)
        B, Tmax = em.shape[:2]
        ends = self._ends(B, vc, addl_ends)
        w = np.full(B, 1.0 / B)
        logz, acc = O.batch_logz_and_counts(em.detach().numpy(), lengths, init.detach().numpy(),
                                            trans.detach().numpy(), lenp.detach().numpy(), ends, w)
        Kc = acc["E_len"].shape[0]
        surrogate = (torch.tensor(acc["E_init"]) * init).sum() + (torch.tensor(acc["E_trans"]) * trans).sum() \
            + (torch.tensor(acc["E_len"]) * lenp[:Kc]).sum() + (torch.tensor(acc["E_em"]) * em).sum()
        surrogate.backward()
        grads = {k: (v.grad.numpy() if v.grad is not None else np.zeros(v.shape)) for k, v in p.items() if v.requires_grad}
        return dict(ll=float(logz.mean()), logz=logz, grads=grads, elp=em.detach().numpy(), counts=acc,
                    init=init.detach().numpy(), trans=trans.detach().numpy(), lenp=lenp.detach().numpy(), ends=ends)

    def viterbi(self, features, lengths, valid=None, addl_ends=None, constraints=None):
        """Span-encoded predictions in GLOBAL class ids (EOS -> n_classes), as the reference returns."""
        p, vc, init, trans, lenp, em = self._scores(features, valid, constraints)
        B = em.shape[0]
        ends = self._ends(B, vc, addl_ends)
        spans, best = O.batch_viterbi(em.detach().numpy(), lengths, init.detach().numpy(),
                                      trans.detach().numpy(), lenp.detach().numpy(), ends)
        table = np.concatenate([vc.numpy(), [self.n_classes]])
        out = np.where(spans >= 0, table[np.clip(spans, 0, None)], -1)
        return out, best, dict(em=em.detach().numpy(), init=init.detach().numpy(), trans=trans.detach().numpy(),
                               lenp=lenp.detach().numpy(), ends=ends, local_spans=spans)


def from_golden(g):
    """Build a ModuleOracle from a tests/golden/*.npz fixture."""
    merge = None
    if "merge_src" in g:
        merge = {int(s): int(d) for s, d in zip(g["merge_src"], g["merge_dst"])}
    return ModuleOracle(
        {k: g[k] for k in ("gaussian_means", "gaussian_cov", "transition_logits", "init_logits", "poisson_log_rates")},
        int(g["max_k"]),
        init_constraints=g["init_constraints"] if "init_constraints" in g else None,
        transition_constraints=g["transition_constraints"] if "transition_constraints" in g else None,
        allowed_ends=g["allowed_ends"] if "allowed_ends" in g else None,
        merge_classes=merge,
    )


def golden_addl_ends(g):
    if "addl_ends_flat" not in g:
        return None
    return [[] if int(x) < 0 else [int(x)] for x in g["addl_ends_flat"]]
