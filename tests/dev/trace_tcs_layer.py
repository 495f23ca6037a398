"""Developer trace of the fused fp16x2 layer kernel (not a test).  Build a trace library first:
   make -C wavenet_b200/csrc clean all EXTRA=-DWN_LAYER_TRACE OUT=../libwavenet_b200_trace.so ; WN_LIB_PATH=... python this"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from bench import config_c, synth_batch
from wavenet_b200 import _lib
from wavenet_b200.wavenet import WaveNet, _stream

lib = _lib.load()
B, W = 32, 16000
net = WaveNet(config_c(), seed=0)
net.to_gpu(0)
net.set_precision("fp16x2")
net.update_laerning_rate(1e-3)
x, t = synth_batch(0, B, W)
xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
net.use_cuda_graph = False
net.train_step(xd, td)
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for _ in range(3):
    _lib.check(lib.wn_tcs_layer_forward(net._h, layer, _stream()))
torch.cuda.synchronize()
buf = np.zeros(64 * 32, dtype=np.int64)
fn = lib.wn_debug_tcs_trace
fn.argtypes = [ctypes.c_void_p]
fn.restype = ctypes.c_int
assert fn(buf.ctypes.data) == 0
tr = buf.reshape(64, 32)
t0 = tr[0][2]
names = {0: "P d2_full", 1: "P rd_done", 2: "P loaded", 3: "P z_full", 4: "M a_full", 5: "M g1 issued", 6: "E g2 issue",
         8: "E d1", 9: "E zarr", 10: "E d2", 11: "E xo_stored"}
for j in range(4, 12):
    print("tile", j, " ".join("%s=%d" % (names[e], tr[j][e] - t0) for e in sorted(names) if tr[j][e]))
print("period", (tr[20][8] - tr[4][8]) / 16)
