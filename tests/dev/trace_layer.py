"""Developer trace of the fused layer kernel (not a test).  Needs `make -C wavenet_b200/csrc clean all EXTRA=-DWN_LAYER_TRACE`."""
import sys, ctypes
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
for l in (12, 12, 12):
    _lib.check(lib.wn_tc_layer_forward(net._h, l, _stream()))
torch.cuda.synchronize()
buf = np.zeros(64 * 32, dtype=np.int64)
raw = ctypes.CDLL(lib._name) if hasattr(lib, "_name") else lib
fn = raw.wn_debug_layer_trace; fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
assert fn(buf.ctypes.data) == 0
tr = buf.reshape(64, 32)
t0 = tr[0][2]
names = {0:"P d2_full", 1:"P rd_done", 2:"P loaded", 3:"P z_full", 4:"M a_full", 5:"M g1 issued", 6:"M z_full",
         8:"E d1", 9:"E zarr", 10:"E d2", 11:"E xo_stored"}
for j in range(4, 12):
    print("tile", j, " ".join("%s=%d" % (names[e], tr[j][e] - t0) for e in sorted(names) if tr[j][e]))
print("period", (tr[20][8] - tr[4][8]) / 16)
