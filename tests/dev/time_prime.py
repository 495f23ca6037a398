"""Developer timing: FasterWaveNet.prime (fp32 SIMT full pass, slices of 256 streams) by stream count."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench import config_c
from wavenet_b200.faster_wavenet import FasterWaveNet
for n in [int(v) for v in sys.argv[1:]] or [256, 1920]:
    net = FasterWaveNet(config_c(), seed=1234)
    net.to_gpu(0)
    window = np.random.default_rng(0).integers(0, 256, (n, net.input_width)).astype(np.int32)
    net.prime(window); torch.cuda.synchronize()
    t0 = time.perf_counter(); net.prime(window); torch.cuda.synchronize()
    print("n=%d prime %.3f s" % (n, time.perf_counter() - t0), flush=True)
    del net; torch.cuda.empty_cache()
