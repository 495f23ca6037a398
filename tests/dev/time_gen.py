"""Developer timing of the generator kernels: streams per GPU x kernel choice.  usage: python tests/dev/time_gen.py [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import config_c
from wavenet_b200 import _lib
from wavenet_b200.faster_wavenet import FasterWaveNet
from wavenet_b200.wavenet import _ptr, _stream

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for n, env in ((1, {}), (8, {}), (256, {}), (32, {"WN_GEN_V5": "1"})):
    for k in ("WN_GEN_V5",):
        os.environ.pop(k, None)
    os.environ.update(env)
    net = FasterWaveNet(config_c(), seed=1234)
    net.to_gpu(0)
    window = np.random.default_rng(0).integers(0, 256, (n, net.input_width)).astype(np.int32)
    net.prime(window)
    out = torch.empty((n, steps), dtype=torch.int32, device="cuda")
    run = lambda: _lib.check(net._libh.wn_gen_run(net._gen, _ptr(net._params), steps, _lib.WN_GEN_SAMPLE, 0, _ptr(out), _stream()))
    run()
    torch.cuda.synchronize()
    net.prime(window)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("n=%3d %-18s %7.2f us/step  %8.3f M samples/s" % (n, str(env), 1e3 * ms / steps, n * steps / ms / 1e3), flush=True)
    del net
