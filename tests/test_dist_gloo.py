"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: batch-sharded gradients averaged by one
all-reduce equal the full-batch gradient; generation streams shard without overlap."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import wavenet_oracle as O
from tests.util import make_cfg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from wavenet_b200.dist import shard_range, allreduce_sum_, assert_replicas_equal
    cfg = make_cfg("tiny_k2")
    w = O.init_weights(cfg, np.random.default_rng(0), np.float64)
    rng = np.random.default_rng(1)
    x = rng.integers(0, 6, (4, 30))
    t = rng.integers(0, 6, (4, 30))
    b0, b1 = shard_range(4, world, rank)
    g = O.backward(cfg, O.forward_loss(cfg, w, x[b0:b1], t[b0:b1], dtype=np.float64))
    names = [n for n, _ in O.param_shapes(cfg)]
    flat = torch.from_numpy(np.concatenate([g[n].reshape(-1) for n in names]))
    scale = allreduce_sum_(flat)
    flat *= scale
    g_full = O.backward(cfg, O.forward_loss(cfg, w, x, t, dtype=np.float64))
    want = np.concatenate([g_full[n].reshape(-1) for n in names])
    ok_grad = np.abs(flat.numpy() - want).max() < 1e-12
    ok_rep = assert_replicas_equal(flat.clone(), atol=0.0)
    lo, hi = shard_range(257, world, rank)
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([lo, hi]))
    ok_shard = sizes[0][0].item() == 0 and sizes[-1][1].item() == 257 and all(
        sizes[i][1].item() == sizes[i + 1][0].item() for i in range(world - 1))
    ret[rank] = (bool(ok_grad), bool(ok_rep), bool(ok_shard), scale)
    dist.destroy_process_group()


def test_dp_gradient_average_and_stream_sharding_world2():
    world = 2
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r] == (True, True, True, 0.5), ret[r]


def test_shard_range_covers_everything():
    from wavenet_b200.dist import shard_range
    for n in (1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            got = [shard_range(n, world, r) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in got) - min(b - a for a, b in got) <= 1
