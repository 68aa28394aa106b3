"""How far is the REFERENCE's own fp32 algorithm from the fp64 oracle at a BASELINE-size video?

Runs the unmodified `SemiMarkovModule.log_hsmm` (materialised potentials) + the pytorch-struct DP restated in
oracle/torch_struct_shim.py + autograd marginals in float32 on one chain-constrained video (T = 3000, C = 23, K = 20,
the configs[1] shape) and prints the deviation of logZ and of the four expected-count tensors from
oracle/hsmm_oracle.py (float64).  ~75 s on 8 cores.  Output committed as profiles/r02_reference_fp32_noise_T3000.txt;
tests/test_gpu_general_fullsize.py cites it for its full-size tolerance."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from oracle import hsmm_oracle as O
from oracle.torch_struct_shim import SemiMarkovCRF
from tests.helpers import random_problem
from tests.golden.ref_import import load_reference
mods, utils = load_reference()
rng = np.random.default_rng(33)
B, Tmax, C, K = 1, 3000, 23, 20
prob = random_problem(rng, B, Tmax, C, K, Tmin=3000, chain=True, ends=True, scale=2.5)
f32 = lambda x: x.astype(np.float32).astype(np.float64)
em, init, trans, lenp, end = f32(prob["em"]), f32(prob["init"]), f32(prob["trans"]), f32(prob["lenp"]), prob["end"]
ref_logz, acc = O.batch_logz_and_counts(em, prob["lengths"], init, trans, lenp, end, np.ones(B))
t = lambda x: torch.tensor(x, dtype=torch.float32)
emt = t(em).requires_grad_(True); tr = t(trans).requires_grad_(True); it = t(init).requires_grad_(True); ln = t(lenp).requires_grad_(True)
lengths = torch.LongTensor(prob["lengths"])
t0=time.time()
scores = mods.SemiMarkovModule.log_hsmm(tr, emt, it, ln, lengths, add_eos=True, allowed_ends_per_instance=[[C-1]])
dist = SemiMarkovCRF(scores, lengths=lengths+1)
lz = dist.partition
lz.sum().backward()
print("time", time.time()-t0)
print("logz err", float(lz[0]) - ref_logz[0])
for name, g, r in [("E_em", emt.grad.numpy(), acc["E_em"]), ("E_trans", tr.grad.numpy(), acc["E_trans"]), ("E_len", ln.grad.numpy(), acc["E_len"][:ln.shape[0]]), ("E_init", it.grad.numpy(), acc["E_init"])]:
    print(name, np.abs(g - r).max() / np.abs(r).max())
print("rowsum-1", np.abs(emt.grad.numpy()[0].sum(axis=1) - 1).max())
