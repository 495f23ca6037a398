"""Developer trace of gen_kernel_v6 (not a test).  Needs a library built with EXTRA=-DWN_LAYER_TRACE (WN_LIB_PATH)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from bench import config_c
from wavenet_b200 import _lib
from wavenet_b200.faster_wavenet import FasterWaveNet

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
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
print("step start -> x0 written (sampling exchange, embedding):", tr[41][1] - t0)
print("epilogue thread 0: start | gate-done wait | gate epi + z send | proj-done wait | x epilogue ||"
      " MMA thread (rel. to layer start): weights, x(t-d), x(t) ready, z ready, issued || producer: x ready, xd loaded, store read, xd issued, w issued")
for l in (0, 1, 2, 3, 10, 11, 20, 28, 29):
    r = tr[l]
    print("layer %2d: %6d | %5d %5d %5d %5d || %5d %5d %5d %5d %5d || %5d %5d %5d %5d %5d" %
          (l, r[0] - t0, r[1] - r[0], r[2] - r[1], r[3] - r[2], r[4] - r[3],
           r[5] - r[0], r[6] - r[0], r[7] - r[0], r[8] - r[0], r[9] - r[0],
           r[10] - r[0], r[11] - r[0], r[12] - r[0], r[13] - r[0], r[14] - r[0]))
print("sampling: candidates ready %d | exchange %d | embedding %d | x0 stored %d" % (tr[42][0] - t0, tr[42][1] - tr[42][0], tr[42][2] - tr[42][1], tr[41][1] - tr[42][2]))
h0 = tr[40][0]
print("head: h0 sent %d | conv0 done %d | h1 sent %d | conv1 done %d | end %d" % (tr[43][0] - h0, tr[43][1] - h0, tr[43][2] - h0, tr[43][3] - h0, tr[40][1] - h0))
print("head detail (rel. head start): conv0: stored %d fenced %d synced %d sent %d | conv1: stored %d fenced %d synced %d hfree %d" % tuple(tr[44][i] - h0 for i in range(8)))
print("sampling detail (rel. step start): fenced %d synced %d | scan done %d" % (tr[45][0] - t0, tr[45][1] - t0, tr[45][2] - t0))
print("layers total:", tr[40][0] - tr[41][1], " head:", tr[40][1] - tr[40][0], " step:", tr[40][1] - t0)
