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
    """In-place sum over ranks of a flat buffer through torch.distributed (host-side logic / CPU tests); returns the
    scale that turns it into the mean.  The GPU training path uses the communicator behind the C ABI instead
    (init_comm / wn_allreduce_grads)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


def init_comm(net, rank=None, world=None, exchange=None):
    """Creates the NCCL communicator INSIDE libwavenet_b200.so for `net` (wn_comm_init): rank 0 makes the 128-byte id,
    `exchange(id_bytes_or_None) -> id_bytes` ships it to every rank (default: torch.distributed.broadcast_object_list,
    which only carries these 128 bytes -- the gradient all-reduce itself never goes through torch)."""
    import ctypes as C
    from ._lib import check
    if rank is None:
        rank, world = dist.get_rank(), dist.get_world_size()
    buf = C.create_string_buffer(128)
    if rank == 0:
        check(net._libh.wn_comm_unique_id(buf))
    if exchange is None:
        box = [bytes(buf.raw) if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ident = box[0]
    else:
        ident = exchange(bytes(buf.raw) if rank == 0 else None)
    idbuf = C.create_string_buffer(ident, 128)
    torch.cuda.set_device(net._device)
    check(net._libh.wn_comm_init(net._h, idbuf, int(rank), int(world)))
    net.data_parallel = world > 1
    return world


def assert_replicas_equal(flat, atol=0.0):
    """Every rank must hold identical parameters after an update (deterministic replicated Adam)."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return True
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    ok = torch.tensor([1 if (flat - ref).abs().max().item() <= atol else 0], device=flat.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    return bool(ok.item())
