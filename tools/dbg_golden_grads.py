import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')

from tests.helpers import *
from oracle.module_oracle import from_golden, golden_addl_ends
import numpy as np, os
def golden(name):
    return dict(np.load(os.path.join('tests/golden', name + '.npz'), allow_pickle=True))
KEYMAP = dict(g_means="gaussian_means", g_trans="transition_logits", g_init="init_logits", g_rates="poisson_log_rates")
for case in ["unconstrained", "short_clamp", "constrained", "constrained_narration"]:
    g = golden(case)
    m = module_from_golden(g)
    feats = torch.from_numpy(g["features"]).cuda(); lengths = torch.from_numpy(g["lengths"]).long()
    B = feats.shape[0]
    vpi = [torch.from_numpy(g["valid_classes"]).long() for _ in range(B)] if "valid_classes" in g else None
    cons = torch.from_numpy(g["constraints"]).cuda() if "constraints" in g else None
    addl = golden_addl_ends(g)
    ll, _ = m.log_likelihood(feats, lengths, vpi, spans=None, add_eos=True, additional_allowed_ends_per_instance=addl, constraints=cons)
    ll.backward()
    r = from_golden(g).log_likelihood(g["features"], g["lengths"], g.get("valid_classes"), addl, g.get("constraints"))
    print(case, float(ll), float(g["ll"]), r["ll"])
    for gk, pk in KEYMAP.items():
        mine = getattr(m, pk).grad.cpu().numpy()
        print("  ", gk, "mine-vs-oracle %.3e  ref-vs-oracle %.3e  mine-vs-ref %.3e" % (rel_err(mine, r["grads"][pk]), rel_err(g[gk], r["grads"][pk]), rel_err(mine, g[gk])))
        if gk in ("g_trans",) and case == "constrained_narration":
            np.set_printoptions(precision=8, suppress=True, linewidth=220)
            d = mine - r["grads"][pk]
            i, j = np.unravel_index(np.abs(d).argmax(), d.shape)
            print("   worst", i, j, mine[i, j], r["grads"][pk][i, j], g[gk][i, j])
            print("   mine  ", mine[:5, :5].ravel())
            print("   oracle", r["grads"][pk][:5, :5].ravel())
            print("   ref   ", g[gk][:5, :5].ravel())
    if case == "constrained_narration":
        print("logz golden", g["logz"])
