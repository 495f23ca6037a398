"""Deterministic mode at the full size (config C, 32 x 16000): two runs of three train steps in each mode -- are the weights
bit-identical? -- and the step time of both modes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from wavenet_b200.wavenet import WaveNet

B, W = 32, 16000
x_h, t_h = bench.synth_batch(0, B, W)
x_d, t_d = torch.from_numpy(x_h).cuda(), torch.from_numpy(t_h).cuda()


def run(det, steps=3, time_it=False):
    net = WaveNet(bench.config_c(), seed=1234)
    net.to_gpu(0)
    net.set_precision("fp16x2")
    net.update_laerning_rate(1e-3)
    if det:
        net.set_deterministic(True)
    for _ in range(steps):
        net.train_step(x_d, t_d)
    torch.cuda.synchronize()
    ms = None
    if time_it:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            net.train_step(x_d, t_d)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
    flat = net._params.detach().cpu().numpy().copy()
    del net
    torch.cuda.empty_cache()
    return flat, ms


for det in (False, True):
    a, ms = run(det, time_it=True)
    b, _ = run(det, steps=8)
    c, _ = run(det, steps=8)
    print("deterministic=%s: %.3f ms/step; two runs of 8 steps bit-identical: %s (max |diff| %.3e, %d of %d weights differ)" % (
        det, ms, np.array_equal(b, c), np.abs(b - c).max(), int((b != c).sum()), b.size), flush=True)
