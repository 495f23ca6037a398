"""Developer experiment: is the error of a long-K split GEMM dominated by in-tensor-core accumulation (round toward zero)?
One K=1920 GEMM vs 30 K=64 GEMMs summed in fp32 outside, both against fp64."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from wavenet_b200 import _lib
from tests.util import make_cfg, make_net
from oracle import wavenet_oracle as O

lib = _lib.load()
P = C.c_void_p
lib.wn_tcs_debug_gemm.restype = C.c_int
lib.wn_tcs_debug_gemm.argtypes = [P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P, C.c_int, P, P, C.c_int, P,
                                  C.c_int, P]


def split(x, scale=1.0):
    x = x * scale
    hi = x.half()
    lo = (x - hi.float()).half()
    return torch.cat([hi, lo], dim=-1).contiguous()


def unsplit(s):
    c = s.shape[-1] // 2
    return s[..., :c].double() + s[..., c:].double()


def ptr(t):
    return P(0) if t is None else P(t.data_ptr())


cfg = make_cfg("C_small")
net = make_net(cfg, O.init_weights(cfg, np.random.default_rng(0), np.float64))
h = net._h
st = P(torch.cuda.current_stream().cuda_stream)
g = torch.Generator(device="cuda").manual_seed(1)
rows, N = 2048, 256
for label, K, pos in [("K=1920 zero-mean", 1920, False), ("K=1920 A>=0,W>=0", 1920, True), ("K=128 zero-mean", 128, False)]:
    A = torch.randn(1, rows, K, device="cuda", generator=g)
    Wt = torch.randn(N, K, device="cuda", generator=g) / np.sqrt(K)
    if pos:
        A, Wt = A.abs(), Wt.abs()
    As, Ws = split(A, 8.0), split(Wt, 128.0)
    ref = (unsplit(As)[0] @ unsplit(Ws).T) / 1024.0
    Y = torch.zeros(1, rows, N, device="cuda")
    assert lib.wn_tcs_debug_gemm(h, ptr(As), K, rows, 1, 1, 0, 0, rows, ptr(Ws), N, None, None, 0, ptr(Y), 0, st) == 0
    torch.cuda.synchronize()
    one = Y[0].double() / 1024.0
    acc = torch.zeros(rows, N, device="cuda")
    for kb in range(K // 64):
        Ak = torch.cat([As[..., kb * 64:(kb + 1) * 64], As[..., K + kb * 64:K + (kb + 1) * 64]], dim=-1).contiguous()
        Wk = torch.cat([Ws[:, kb * 64:(kb + 1) * 64], Ws[:, K + kb * 64:K + (kb + 1) * 64]], dim=-1).contiguous()
        Yk = torch.zeros(1, rows, N, device="cuda")
        assert lib.wn_tcs_debug_gemm(h, ptr(Ak), 64, rows, 1, 1, 0, 0, rows, ptr(Wk), N, None, None, 0, ptr(Yk), 0, st) == 0
        acc += Yk[0]
    torch.cuda.synchronize()
    chunked = acc.double() / 1024.0
    f32 = ((A[0] * 8.0) @ (Wt * 128.0).T).double() / 1024.0     # torch fp32 matmul (no tf32) for comparison
    sc = ref.abs().max().item()
    print("%s: |ref|max %.2f  one-GEMM err %.2e (mean signed %.2e)  chunked-64 err %.2e (mean signed %.2e)  torch-fp32 err %.2e" %
          (label, sc, (one - ref).abs().max().item() / sc, ((one - ref) * ref.sign()).mean().item() / sc,
           (chunked - ref).abs().max().item() / sc, ((chunked - ref) * ref.sign()).mean().item() / sc,
           (f32 - ref).abs().max().item() / sc))
