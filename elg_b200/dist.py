"""Multi-GPU layout of the inference path: instances are independent, so each rank owns a
contiguous block of instances (with all 8 augmentations and all POMO rows of an instance on the same GPU,
best-of-augmentation needs no exchange) and there is NO collective on the data path; only the per-instance
costs are gathered at the end.  Works with any torch.distributed backend (nccl on GPUs, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def shard_slice(n, rank, world):
    """Contiguous, balanced block of range(n) owned by `rank`: sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def shard_batch(batch, rank, world):
    """Slice every tensor of a batch (dict or tensor) along dim 0 for this rank."""
    n = (next(iter(batch.values())) if isinstance(batch, dict) else batch).shape[0]
    s = shard_slice(n, rank, world)
    return {k: v[s] for k, v in batch.items()} if isinstance(batch, dict) else batch[s]


def gather_costs(local_costs, n_total):
    """All ranks contribute their block of per-instance costs; every rank gets the (n_total,) vector in
    instance order.  The only communication of the inference path (a few bytes per instance)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_costs
    world, rank = dist.get_world_size(), dist.get_rank()
    base, rem = divmod(n_total, world)
    pad = base + (1 if rem else 0)
    buf = torch.zeros(pad, dtype=local_costs.dtype, device=local_costs.device)
    buf[:local_costs.numel()] = local_costs
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return torch.cat([out[r][:shard_slice(n_total, r, world).stop - shard_slice(n_total, r, world).start] for r in range(world)])


def allreduce_mean_gradient(flat_grad, group=None):
    """Training path: sum the packed gradient over the ranks (ONE all-reduce of the whole parameter vector, 5 MB for the
    released model) and return the factor that turns the sum into the gradient of the global mean of J.  Every rank's J
    is a mean over its own instances x POMO rows (CVRP/train.py:121), so with equal shards mean-of-means = global mean
    (SURVEY 8e); the factor is applied inside elg_adam_step (grad_scale)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / dist.get_world_size(group)
