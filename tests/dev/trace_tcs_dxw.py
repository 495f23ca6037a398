"""Developer trace of tcs_dxw_kernel (not a test): the LAST launch of a train step (layer 0).  Needs a library built with
EXTRA=-DWN_LAYER_TRACE (WN_LIB_PATH=...)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from bench import config_c, synth_batch
from wavenet_b200 import _lib
from wavenet_b200.wavenet import WaveNet

lib = _lib.load()
B, W = 32, 16000
net = WaveNet(config_c(), seed=0)
net.to_gpu(0)
net.set_precision("fp16x2")
net.update_laerning_rate(1e-3)
x, t = synth_batch(0, B, W)
xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
net.use_cuda_graph = False
for _ in range(2):
    net.train_step(xd, td)
torch.cuda.synchronize()
buf = np.zeros(64 * 32, dtype=np.int64)
fn = lib.wn_debug_tcs_trace
fn.argtypes = [ctypes.c_void_p]
fn.restype = ctypes.c_int
assert fn(buf.ctypes.data) == 0
tr = buf.reshape(64, 32)
t0 = tr[4][0]
names = ["P u0", "P u1", "P u2", "P u3", "M u0", "M u1", "M u2", "M u3", "M done", "E acc", "E stored", "", "P xwait", "P xfree", "M xwait", "M xfull"]
for j in range(4, 12):
    print("tile %2d " % j + " ".join("%s=%d" % (names[e], tr[j][e] - t0) for e in (0, 1, 12, 13, 2, 3, 4, 5, 14, 15, 6, 7, 8, 9, 10)))
print("period", (tr[20][8] - tr[4][8]) / 16)
