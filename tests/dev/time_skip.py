import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from wavenet_b200 import _lib
from wavenet_b200.wavenet import WaveNet, _ptr, _stream
lib = _lib.load()
B, W = 32, 16000
net = WaveNet(bench.config_c(), seed=1234); net.to_gpu(0); net.set_precision("fp16x2")
x, t = bench.synth_batch(0, B, W)
xd, td = torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda()
net._bind(B, W)
_lib.check(lib.wn_forward_loss(net._h, _ptr(net._params), _ptr(xd), _ptr(td), W, _ptr(net._loss), None, _stream()))
for name, fn in (("skip gemm", lambda: _lib.check(lib.wn_tcs_skip_gemm(net._h, _stream()))),):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.1f us" % (1e3 * e0.elapsed_time(e1) / 10))
