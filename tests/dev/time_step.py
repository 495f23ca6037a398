"""Developer timing: config-C train step (32 x 16000) per precision mode, CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench

precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tf32", "fp16x2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, W = 32, 16000
from wavenet_b200.wavenet import WaveNet
x_h, t_h = bench.synth_batch(0, B, W)
x_d, t_d = torch.from_numpy(x_h).cuda(), torch.from_numpy(t_h).cuda()
for prec in precs:
    net = WaveNet(bench.config_c(), seed=1234)
    net.to_gpu(0)
    net.set_precision(prec)
    net.update_laerning_rate(1e-3)
    for _ in range(3):
        net.train_step(x_d, t_d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = net.train_step(x_d, t_d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("%s: %.3f ms/step  %.2f M samples/s  loss %.4f" % (prec, ms, B * W / ms / 1e3, float(loss[0])), flush=True)
    del net
    torch.cuda.empty_cache()
