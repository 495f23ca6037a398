"""Small workload for compute-sanitizer (racecheck / synccheck / memcheck): one fp16x2 train step of the fused-shape
network (tcs_layer / tcs_gemm incl. the fused CE epilogue / tcs_gate_bwd / tcs_dxw / tcs_wgrad), one tf32 step, and a few
steps of the cluster generator (gen_kernel_v4) and of the single-CTA generator (gen_kernel_v3)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
if what in ("all", "train"):
    cfg = make_cfg("C_small")
    w = O.init_weights(cfg, np.random.default_rng(0), np.float64)
    rng = np.random.default_rng(1)
    x = rng.integers(0, 256, (2, 300)).astype(np.int32)
    t = rng.integers(0, 256, (2, 300)).astype(np.int32)
    for prec in ("fp16x2", "fp16x2-deterministic", "tf32"):
        net = make_net(cfg, w)
        net.set_precision(prec.split("-")[0])
        net.use_cuda_graph = False
        net.update_laerning_rate(1e-3)
        if prec.endswith("deterministic"):
            net.set_deterministic(True)
        loss = net.train_step(dev(x), dev(t))
        torch.cuda.synchronize()
        print("train", prec, float(loss[0]), flush=True)
if what in ("all", "gen"):
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(0), np.float32)
    for n, tag in ((1, "v4 cluster"), (20, "v3 single CTA")):
        net = make_net(cfg, w, faster=True)
        win = np.random.default_rng(2).integers(0, 256, (n, O.input_width(cfg))).astype(np.int32)
        out = net.generate(win, 6, mode="greedy")
        torch.cuda.synchronize()
        print("gen", tag, out[0].tolist(), flush=True)
if what in ("gen6", "gen6_4"):
    # gen_kernel_v6 (tensor-core generator): one partially filled cluster, forced below its automatic threshold
    os.environ["WN_GEN_V6"] = "1"
    os.environ["WN_GEN_V6_CS"] = "4" if what == "gen6_4" else "8"
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(0), np.float32)
    net = make_net(cfg, w, faster=True)
    win = np.random.default_rng(2).integers(0, 256, (20, O.input_width(cfg))).astype(np.int32)
    out = net.generate(win, 9, mode="sample")
    torch.cuda.synchronize()
    print("gen v6 tensor-core", out[0].tolist(), flush=True)
