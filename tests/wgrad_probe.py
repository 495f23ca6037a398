"""Manual probe of the MN-major wgrad kernel descriptors (not a pytest)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wavenet_b200 import _lib
from wavenet_b200.wavenet import WaveNet, Params, _ptr, _stream
lib = _lib.load()
lib.wn_debug_wgrad.restype = C.c_int
lib.wn_debug_wgrad.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
net = WaveNet(Params(), seed=0); net.to_gpu()
torch.manual_seed(0)
rows, nseq = 96, 1
for Kd, Kx in ((128, 64),(128,128),(128,256)):
    dY = torch.randn(nseq * rows, Kd, device="cuda")
    X = torch.randn(nseq * rows, Kx, device="cuda")
    ref = (dY[:, :128].double().T @ X.double()).float()
    for (lbo, sbo, kstep, major) in [(0, 0, 0, 0), (4096, 1024, 1024, 0), (4096, 256, 1024, 0), (512, 4096, 1024, 0)]:
        dW = torch.zeros(128, Kx, device="cuda")
        rc = lib.wn_debug_wgrad(net._h, _ptr(dY), Kd, _ptr(X), Kx, rows, nseq, _ptr(dW), lbo, sbo, kstep, major, _stream())
        torch.cuda.synchronize()
        err = (dW - ref).abs().max().item()
        print("Kd=%d Kx=%d lbo=%d sbo=%d kstep=%d major=%d rc=%d  |dW|=%.3e |ref|=%.3e maxerr=%.3e  dW[0,:4]=%s ref[0,:4]=%s" % (
            Kd, Kx, lbo, sbo, kstep, major, rc, dW.norm().item(), ref.norm().item(), err, dW[0, :4].tolist(), ref[0, :4].tolist()), flush=True)
        print("   dW[1,:6]", dW[1,:6].tolist(), " dY[0,:6]", dY[0,:6].tolist(), " dY[1,:3]", dY[1,:3].tolist(), "X[0,:4]", X[0,:4].tolist())
