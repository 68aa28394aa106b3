"""Host logic of `SemiMarkovModel` on CPU (no kernels): pickling, and the data-parallel `fit` under gloo with
world_size 2 and a mini-batch SMALLER than the world (rank 1 owns no video of any batch: it must contribute zero
gradients and still join every all-reduce), against a single-process run.

The DP itself cannot run here (no GPU, no CPU fallback in the product): the module's `log_likelihood` is replaced, in
this test only, by the fp64 oracle wrapped as an autograd surrogate -- the checker standing in for the kernels so that
the trainer's sharding / all-reduce / optimiser plumbing can be exercised."""
import os
import pickle

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import action_segmentation_b200 as pkg
from action_segmentation_b200 import data, semimarkov
from action_segmentation_b200.args import HsmmArgs


def _oracle_backed(model):
    """Replace the CUDA-only methods of `model.model` by oracle-backed CPU stand-ins (test only)."""
    from oracle.module_oracle import ModuleOracle
    m = model.model

    def log_likelihood(features, lengths, valid_classes_per_instance, spans=None, add_eos=True, use_mean_z=False,
                       additional_allowed_ends_per_instance=None, constraints=None):
        params = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        mo = ModuleOracle(params, m.max_k, init_constraints=params.get("init_constraints"),
                          transition_constraints=params.get("transition_constraints"), allowed_ends=m.allowed_ends,
                          merge_classes=m.merge_classes)
        valid = None if valid_classes_per_instance is None else valid_classes_per_instance[0].numpy()
        r = mo.log_likelihood(features.numpy(), lengths.numpy(), valid, additional_allowed_ends_per_instance,
                              None if constraints is None else constraints.numpy())
        ll = torch.tensor(r["ll"], dtype=torch.float32)
        for k, g in r["grads"].items():
            p = getattr(m, k)
            ll = ll + (torch.from_numpy(g).float() * (p - p.detach())).sum()
        m.kl = torch.zeros(features.size(0))
        return ll, torch.zeros(())

    def initialize_gaussian(feats, lengths):
        x = torch.cat([feats[i, :int(lengths[i])] for i in range(feats.size(0))], dim=0)
        m.gaussian_means.data.copy_(x.mean(dim=0, keepdim=True).expand_as(m.gaussian_means))
        m.gaussian_cov.data = torch.diag(x.var(dim=0))

    m.log_likelihood = log_likelihood
    m.initialize_gaussian = initialize_gaussian


def _train(rank, world, port, out):
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    semimarkov.DeviceBatchCache.to_device = classmethod(lambda cls, batch: dict(batch))  # no GPU here
    torch.Tensor.cuda = lambda self, *a, **k: self
    split = data.make_crosstask_like(n_tasks=2, steps_per_task=(2, 2), n_videos=5, feature_dim=4, frames=(14, 20), seed=2)
    args = HsmmArgs(sm_max_span_length=6, sm_constrain_transitions=True, annotate_background_with_previous=True, epochs=2,
                    batch_size=1, training='unsupervised', print_every=0, lr=0.05)
    torch.manual_seed(100 + rank)  # replicas start DIFFERENT: fit() must broadcast rank 0's parameters
    model = pkg.SemiMarkovModel.from_args(args, split)
    if world > 1 and rank == 0 or world == 1:
        torch.manual_seed(100)
        torch.nn.init.uniform_(model.model.init_logits, 0, 1)
    _oracle_backed(model)
    log = []
    model.fit(split, use_labels=False, callback_fn=lambda e, s: log.append(s['train_loss']))
    out[(world, rank)] = ({k: v.detach().clone() for k, v in model.model.state_dict().items()}, log)
    if world > 1:
        dist.destroy_process_group()


def test_fit_shards_batches_smaller_than_the_world():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_train, args=(2, 29581, out), nprocs=2, join=True)
    _train(0, 1, 0, out)
    single, log1 = out[(1, 0)]
    for rank in range(2):
        state, log = out[(2, rank)]
        assert np.allclose(log, log1, rtol=1e-5), (log, log1)
        for k, v in single.items():
            assert torch.allclose(state[k].float(), v.float(), rtol=1e-4, atol=1e-6), (rank, k)
    assert log1[1] < log1[0]


def test_model_pickle_round_trip_restores_loader_and_cache():
    split = data.make_supervised_like(n_tasks=1, n_videos=3, feature_dim=4, frames=(10, 12))
    model = pkg.SemiMarkovModel.from_args(HsmmArgs(sm_max_span_length=5), split, make_data_loader=data.make_data_loader)
    model._cache = semimarkov.DeviceBatchCache()
    clone = pickle.loads(pickle.dumps(model))
    assert clone._make_data_loader is None and clone._cache is None and clone.dist_group is None
    loader = clone._loader(split, shuffle=False, batch_by_task=True, batch_size=2)  # resolved lazily after unpickling
    assert sum(len(b['lengths']) for b in loader) == 3
    for k, v in model.model.state_dict().items():
        assert torch.equal(clone.model.state_dict()[k], v)
    # pickles written before _cache/_make_data_loader existed still load
    legacy = dict(model.__getstate__())
    legacy.pop('_cache')
    legacy.pop('_make_data_loader')
    m2 = pkg.SemiMarkovModel.__new__(pkg.SemiMarkovModel)
    m2.__setstate__(legacy)
    assert m2._cache is None and m2._loader(split, shuffle=False, batch_by_task=True, batch_size=1) is not None
