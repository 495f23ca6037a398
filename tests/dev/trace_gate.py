"""Developer trace of tc_gate_bwd_kernel (not a test).  Needs `make -C wavenet_b200/csrc clean all EXTRA=-DWN_LAYER_TRACE`."""
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
scratch = torch.zeros_like(net._grads)
for _ in range(3):
    _lib.check(lib.wn_tc_gate_backward_layer(net._h, 12, _ptr(scratch), _stream()))
torch.cuda.synchronize()
buf = np.zeros(64 * 32, dtype=np.int64)
fn = ctypes.CDLL(_lib.LIB_PATH).wn_debug_layer_trace
fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
assert fn(buf.ctypes.data) == 0
tr = buf.reshape(64, 32)
t0 = tr[0][0]
names = {0: "P Kissued", 1: "P mn_empty", 2: "M acc_empty", 3: "M main done", 4: "M mn_full", 5: "M wgrad done",
         8: "E ldg issued", 9: "E acc_full", 10: "E stored"}
for j in range(4, 12):
    print("tile", j, " ".join("%s=%d" % (names[e], tr[j][e] - t0) for e in sorted(names)))
print("period", (tr[20][9] - tr[4][9]) / 16)
