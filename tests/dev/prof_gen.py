import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import wavenet_oracle as O
from bench import config_c
from wavenet_b200.faster_wavenet import FasterWaveNet
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
net = FasterWaveNet(config_c(), seed=0)
net.set_weights(O.init_weights(O.config_C(), np.random.default_rng(1234), np.float32))
net.to_gpu(0)
window = np.random.default_rng(0).integers(0, 256, (n, net.input_width)).astype(np.int32)
out = net.generate(window, steps, mode="sample", seed=0)
torch.cuda.synchronize()
print(out[0, :8].tolist())
