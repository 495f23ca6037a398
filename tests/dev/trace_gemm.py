"""Developer trace of tc_gemm_kernel<256> (not a test).  Needs `make -C wavenet_b200/csrc clean all EXTRA=-DWN_LAYER_TRACE`.
Runs one forward+loss of config C and dumps CTA 0's events of the LAST <256> launch (the second head conv)."""
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
which = sys.argv[1] if len(sys.argv) > 1 else "head"
if which == "skip":
    _lib.check(lib.wn_tc_skip_gemm(net._h, _stream()))
else:
    _lib.check(lib.wn_forward_loss(net._h, _ptr(net._params), _ptr(xd), _ptr(td), W, _ptr(net._loss), None, _stream()))
torch.cuda.synchronize()
buf = np.zeros(64 * 32, dtype=np.int64)
fn = lib.wn_debug_layer_trace if hasattr(lib, "wn_debug_layer_trace") else ctypes.CDLL(_lib.LIB_PATH).wn_debug_layer_trace
fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
assert fn(buf.ctypes.data) == 0
tr = buf.reshape(64, 32)
t0 = tr[0][0]
print("k-step: load issued / full seen")
for it in range(0, 40):
    print(it, tr[it][0] - t0, tr[it][1] - t0)
print("tile: acc_full committed / epilogue start / epilogue end")
for j in range(0, 6):
    print(j, tr[j][2] - t0, tr[j][3] - t0, tr[j][4] - t0)
