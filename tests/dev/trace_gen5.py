"""Developer trace of gen_kernel_v5 (not a test).  Needs a library built with EXTRA=-DWN_LAYER_TRACE (WN_LIB_PATH)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from bench import config_c
from wavenet_b200 import _lib
from wavenet_b200.faster_wavenet import FasterWaveNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = FasterWaveNet(config_c(), seed=1234)
net.to_gpu(0)
window = np.random.default_rng(0).integers(0, 256, (n, net.input_width)).astype(np.int32)
net.generate(window, 20, mode="sample", seed=0)
torch.cuda.synchronize()
buf = np.zeros(64 * 16, dtype=np.int64)
fn = ctypes.CDLL(_lib.LIB_PATH).wn_debug_gen_trace
fn.argtypes = [ctypes.c_void_p]
fn.restype = ctypes.c_int
assert fn(buf.ctypes.data) == 0
tr = buf.reshape(64, 16)
t0 = tr[41][0]
print("step start -> layers start (sampling, taps, embedding):", tr[41][1] - t0)
for l in (0, 1, 2, 10, 11, 28, 29):
    r = tr[l]
    print("layer %2d: start %6d | weights wait %4d | phase A %5d | stage+sync+bulk issue %5d | z wait %5d | phase B %5d | block barrier %4d" %
          (l, r[0] - t0, r[1] - r[0], r[6] - r[1], r[2] - r[6], r[3] - r[2], r[4] - r[3], r[5] - r[4]))
print("layers total:", tr[40][0] - tr[41][1], " head:", tr[40][1] - tr[40][0], " step:", tr[40][1] - t0)
