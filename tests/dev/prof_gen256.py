import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import wavenet_oracle as O
from bench import config_c
from wavenet_b200.faster_wavenet import FasterWaveNet
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
net = FasterWaveNet(config_c(), seed=0)
net.set_weights(O.init_weights(O.config_C(), np.random.default_rng(1234), np.float32))
net.to_gpu(0)
window = np.random.default_rng(0).integers(0, 256, (n, net.input_width)).astype(np.int32)
net.generate(window, 50, mode="sample", seed=0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
net.prime(window)
out = torch.empty((n, steps), dtype=torch.int32, device="cuda")
from wavenet_b200 import _lib
from wavenet_b200.wavenet import _ptr, _stream
e0.record()
_lib.check(net._libh.wn_gen_run(net._gen, _ptr(net._params), steps, 1, 0, _ptr(out), _stream()))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("n=%d NS=%s: %.2f us/step, %.3f M samples/s" % (n, os.environ.get("WN_GEN_NS", "auto"), 1e3 * ms / steps, n * steps / ms / 1e3))
