"""Manual multi-GPU check (run under torchrun, one rank per GPU):
   torchrun --nproc-per-node 2 tests/dp_check.py
 (1) replicas hold identical parameters after data-parallel steps;
 (2) they match a single-process run on the concatenated global batch (gradient mean == global-batch gradient)."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net
from wavenet_b200.dist import assert_replicas_equal

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
cfg = make_cfg("C_small")
w = O.init_weights(cfg, np.random.default_rng(1), np.float32)
B, W = 4, 700
rng = np.random.default_rng(5)
xs = rng.integers(0, 256, (3, B * world, W + 1)).astype(np.int32)
for prec in ("fp32", "tf32"):
    net = make_net(cfg, w); net.set_precision(prec); net.data_parallel = True; net.update_laerning_rate(1e-3)
    for s in range(3):
        x = torch.from_numpy(xs[s, rank * B:(rank + 1) * B]).cuda()
        net.train_step(x[:, :W].contiguous(), x[:, 1:].contiguous())
    ok = assert_replicas_equal(net._params)
    ref = make_net(cfg, w); ref.set_precision(prec); ref.update_laerning_rate(1e-3)
    for s in range(3):
        x = torch.from_numpy(xs[s]).cuda()
        ref.train_step(x[:, :W].contiguous(), x[:, 1:].contiguous())
    diff = (net._params - ref._params).abs().max().item()
    if rank == 0:
        print("%s: replicas identical=%s  max|dp - single|=%.3e" % (prec, ok, diff), flush=True)
    # TF32: atomics + different batch split change rounding; Adam(lr=1e-3) turns tiny gradient differences into +-lr moves
    # fp32: all-reduce summation order differs from the single-process atomics order (6e-5 seen at 8 ranks)
    assert ok and diff < (2e-4 if prec == "fp32" else 5e-3), (ok, diff)
dist.destroy_process_group()
if rank == 0:
    print("DP CHECK OK")
