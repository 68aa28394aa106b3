"""Data-parallel plumbing: videos are independent given the parameters, so a
mini-batch is sharded over ranks and the only exchange is ONE all-reduce (NCCL over NVLink on the
GPU box, gloo in the CPU tests) of a packed fp32 buffer [all parameter gradients | loss]."""
import torch
import torch.distributed as dist


def rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(n, rank, world):
    """Round-robin shard of the n videos of a mini-batch."""
    return list(range(rank, n, world))


def shard_balanced(lengths, rank, world):
    """Longest-processing-time shard: videos sorted by length, dealt to the least loaded rank."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda q: load[q])
        load[r] += int(lengths[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def pack(tensors, extra):
    flat = [t.reshape(-1).to(torch.float32) for t in tensors] + [extra.reshape(-1).to(torch.float32)]
    return torch.cat(flat)


def allreduce_gradients(parameters, loss, group=None):
    """Sum gradients (and the scalar loss) over ranks with a single collective; returns the summed loss."""
    rank, world = rank_world(group)
    if world == 1:
        return loss
    params = [p for p in parameters if p.requires_grad]
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    buf = pack(grads, loss.detach())
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = buf[off:off + n].view_as(g).to(g.dtype).clone()
        off += n
    return buf[off]


def broadcast_parameters(module, group=None, src=0):
    """Make every replica start from rank `src`'s parameters and buffers (one packed broadcast)."""
    rank, world = rank_world(group)
    if world == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n


def allreduce_stats(buf, group=None):
    """In-place SUM of a packed statistics buffer (EM sufficient statistics, frame counters)."""
    rank, world = rank_world(group)
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf
