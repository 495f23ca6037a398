"""ncu helper (not a test): a few forward residual stacks of config C."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import wavenet_oracle as O
from bench import config_c, synth_batch
from wavenet_b200.wavenet import _ptr, _stream
from wavenet_b200 import _lib
from wavenet_b200.faster_wavenet import FasterWaveNet
lib = _lib.load()
B, W = 32, 16000
net = FasterWaveNet(config_c(), seed=0)
net.set_weights(O.init_weights(O.config_C(), np.random.default_rng(1234), np.float32))
net.to_gpu(0); net.set_precision("tf32"); net.update_laerning_rate(1e-3)
x, t = synth_batch(0, B, W)
xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
net.train_step(xd, td)
def layers():
    for l in range(30): _lib.check(lib.wn_tc_layer_forward(net._h, l, _stream()))
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("layer kernel %.1f us" % (1e3 * timed(layers) / 30))
