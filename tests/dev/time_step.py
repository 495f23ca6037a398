"""Quick timing helper (not a test): ms per config-C train step on cuda:0, CUDA events around N graph replays."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from oracle import wavenet_oracle as O
from bench import config_c, synth_batch
from wavenet_b200.faster_wavenet import FasterWaveNet
B, W = 32, 16000
net = FasterWaveNet(config_c(), seed=0)
net.set_weights(O.init_weights(O.config_C(), np.random.default_rng(1234), np.float32))
net.to_gpu(0); net.set_precision("tf32"); net.update_laerning_rate(1e-3)
x, t = synth_batch(0, B, W)
xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
for _ in range(3):
    net.train_step(xd, td)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(n):
    loss = net.train_step(xd, td)
e1.record(); torch.cuda.synchronize()
print("train step %.3f ms  loss %.5f" % (e0.elapsed_time(e1) / n, float(loss)))
