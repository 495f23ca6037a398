"""Host-side helpers for the multi-GPU paths.  Training shards the batch and needs ONE
exchange step (sum all-reduce of the flat gradient, then 1/N inside wn_clip_adam_step); generation shards
independent streams and needs none."""
import torch
import torch.distributed as dist


def shard_range(n_items, world_size, rank):
    """Contiguous [begin, end) of n_items owned by rank; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allreduce_sum_(flat):
    """In-place sum over ranks of the flat gradient buffer; returns the scale that turns it into the mean."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


def assert_replicas_equal(flat, atol=0.0):
    """Every rank must hold identical parameters after an update (deterministic replicated Adam)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return True
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    ok = torch.tensor([1 if (flat - ref).abs().max().item() <= atol else 0], device=flat.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item())
