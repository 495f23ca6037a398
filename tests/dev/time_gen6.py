"""Developer timing: gen_kernel_v6 (tensor-core many-stream generator) against the default kernels by stream count.
usage: python tests/dev/time_gen6.py [steps] [n ...]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from bench import config_c
from wavenet_b200 import _lib
from wavenet_b200.faster_wavenet import FasterWaveNet
from wavenet_b200.wavenet import _ptr, _stream

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ns = [int(v) for v in sys.argv[2:]] or [592, 1024, 2304]
for n in ns:
    for v6 in ("1", "0"):
        os.environ["WN_GEN_V6"] = v6
        try:
            net = FasterWaveNet(config_c(), seed=1234)
            net.to_gpu(0)
            window = np.random.default_rng(0).integers(0, 256, (n, net.input_width)).astype(np.int32)
            net.prime(window)
            out = torch.empty((n, steps), dtype=torch.int32, device="cuda")
            run = lambda: _lib.check(net._libh.wn_gen_run(net._gen, _ptr(net._params), steps, _lib.WN_GEN_SAMPLE, 0, _ptr(out), _stream()))
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            print("n=%4d %s %7.2f us/step  %8.3f M samples/s  (state %.1f GB)" % (n, "v6     " if v6 == "1" else "default", 1e3 * ms / steps, n * steps / ms / 1e3,
                  _lib.load().wn_gen_state_bytes(net._gen) / 1e9), flush=True)
        except Exception as e:
            print("n=%4d v6=%s failed: %s" % (n, v6, str(e)[:200]), flush=True)
        net = run = out = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
