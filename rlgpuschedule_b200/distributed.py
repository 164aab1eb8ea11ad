"""Multi-GPU layout of the path: replicas shard across ranks with no data-path collective; the one
exchange is an all-gather of the per-replica episode returns at the end of a batch
(BASELINE.json north_star).  Backend-agnostic torch.distributed (nccl on GPUs, gloo in
the CPU tests)."""


def shard(n_total, world_size, rank):
    """Contiguous block of replicas owned by `rank`: (first, count).  Blocks differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError('rank %d outside world of %d' % (rank, world_size))
    first = n_total * rank // world_size
    return first, n_total * (rank + 1) // world_size - first


def gather_returns(local_returns, n_total=None, group=None):
    """All-gathers the local int64 returns tensor of every rank (one collective) and returns the
    concatenation in global replica order.  With unequal shards the tensors are padded to the
    largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_returns.clone()
    world = dist.get_world_size(group)
    n_local = local_returns.numel()
    if n_total is None:
        n_total = n_local * world
    counts = [shard(n_total, world, r)[1] for r in range(world)]
    width = max(counts)
    send = local_returns
    if n_local < width:
        send = torch.zeros(width, dtype=local_returns.dtype, device=local_returns.device)
        send[:n_local] = local_returns
    out = torch.empty(world * width, dtype=local_returns.dtype, device=local_returns.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if all(c == width for c in counts):
        return out
    return torch.cat([out[r * width:r * width + counts[r]] for r in range(world)])
