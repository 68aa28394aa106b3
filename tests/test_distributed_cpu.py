"""world_size-2 gloo test of the data-parallel plumbing (one packed all-reduce per step)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from action_segmentation_b200 import distributed as hd
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    frozen = torch.nn.Parameter(torch.ones(2), requires_grad=False)
    x = torch.arange(12, dtype=torch.float32).view(4, 3)
    sel = hd.shard_indices(4, rank, world)
    loss = (lin(x[sel]).sum() ** 2) * (len(sel) / 4.0)
    loss.backward()
    total = hd.allreduce_gradients(list(lin.parameters()) + [frozen], loss.detach())
    stats = hd.allreduce_stats(torch.tensor([float(rank + 1), 1.0]))
    out[rank] = (lin.weight.grad.clone(), lin.bias.grad.clone(), float(total), stats.clone(),
                 hd.shard_balanced([5, 9, 1, 7, 3], rank, world))
    dist.destroy_process_group()


def test_allreduce_matches_single_process():
    world, port = 2, 29571
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    lin = torch.nn.Linear(3, 2)
    x = torch.arange(12, dtype=torch.float32).view(4, 3)
    ref = sum((lin(x[[i for i in range(r, 4, 2)]]).sum() ** 2) * 0.5 for r in range(2))
    ref.backward()
    for r in range(world):
        w, b, total, stats, shard = out[r]
        assert torch.allclose(w, lin.weight.grad) and torch.allclose(b, lin.bias.grad)
        assert abs(total - float(ref)) < 1e-3 * abs(float(ref))
        assert stats.tolist() == [3.0, 2.0]
    assert sorted(out[0][4] + out[1][4]) == [0, 1, 2, 3, 4]
    assert out[0][4] != out[1][4]
