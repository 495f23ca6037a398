"""Developer check of the generic split-fp16 kernels (tcs_gemm_kernel / tcs_wgrad_kernel) against torch fp64 on
caller-made split tensors.  Run on the GPU box:  python tests/dev/check_tcs_kernels.py"""
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
lib.wn_tcs_debug_wgrad.restype = C.c_int
lib.wn_tcs_debug_wgrad.argtypes = [P, P, C.c_int, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, P, P]


def split(x):
    hi = x.half()
    lo = (x - hi.float()).half()
    return torch.cat([hi, lo], dim=-1).contiguous()


def unsplit(s):
    c = s.shape[-1] // 2
    return s[..., :c].float() + s[..., c:].float()


def ptr(t):
    return P(0) if t is None else P(t.data_ptr())


def main():
    cfg = make_cfg("C_small")
    net = make_net(cfg, O.init_weights(cfg, np.random.default_rng(0), np.float64))
    h = net._h
    st = P(torch.cuda.current_stream().cuda_stream)
    g = torch.Generator(device="cuda").manual_seed(1)
    worst = 0.0
    for (K, N, rows, nseq, ns, d, relu, rsd, mask, osplit) in [
            (64, 64, 1000, 2, 1, 0, 0, 0, 0, 0), (64, 128, 1000, 2, 2, 8, 0, 0, 0, 0), (256, 256, 1000, 2, 1, 0, 1, 0, 0, 1),
            (256, 256, 1000, 2, 1, 0, 0, 0, 1, 1), (128, 64, 777, 3, 2, 16, 0, 1, 0, 1), (256, 1920, 300, 1, 1, 0, 0, 0, 0, 0),
            (64, 256, 5000, 5, 1, 0, 0, 0, 0, 1)]:
        A = torch.randn(nseq, rows, K, device="cuda", generator=g)
        Wt = torch.randn(N, ns * K, device="cuda", generator=g) / np.sqrt(ns * K)
        R = torch.randn(nseq, rows, N, device="cuda", generator=g) if rsd else None
        M = torch.randn(nseq, rows, N, device="cuda", generator=g) if mask else None
        As, Ws = split(A), torch.cat([split(Wt)[:, :ns * K], split(Wt)[:, ns * K:]], dim=1).contiguous()
        Rs = split(R) if rsd else None
        Ms = split(M) if mask else None
        Y = torch.zeros(nseq, rows, 2 * N if osplit else N, device="cuda", dtype=torch.float16 if osplit else torch.float32)
        offs = (0, 0) if ns == 1 else (-d, 0)
        rc = lib.wn_tcs_debug_gemm(h, ptr(As), K, rows, nseq, ns, offs[0], offs[1], rows, ptr(Ws), N, ptr(Rs), ptr(Ms), relu, ptr(Y),
                                   osplit, st)
        assert rc == 0, lib.wn_last_error()
        torch.cuda.synchronize()
        got = unsplit(Y).double() if osplit else Y.double()
        A64 = unsplit(As).double()
        ref = torch.zeros(nseq, rows, N, device="cuda", dtype=torch.float64)
        W64 = (Ws[:, :ns * K].float() + Ws[:, ns * K:].float()).double()
        for s_ in range(ns):
            sh = -offs[s_]
            Ash = torch.zeros_like(A64)
            if sh == 0:
                Ash = A64
            else:
                Ash[:, sh:] = A64[:, :-sh]
            ref += Ash @ W64[:, s_ * K:(s_ + 1) * K].T
        if relu:
            ref = ref.clamp_min(0)
        if rsd:
            ref += unsplit(Rs).double()
        if mask:
            ref = ref * (unsplit(Ms) > 0)
        err = ((got - ref).abs().max() / ref.abs().max()).item()
        worst = max(worst, err)
        print("gemm K=%d N=%d rows=%d ns=%d relu=%d rsd=%d mask=%d out_split=%d: max rel-to-max err %.2e" %
              (K, N, rows, ns, relu, rsd, mask, osplit, err))
    for (M_, Cx, rows, nseq, off) in [(64, 64, 1000, 2, 0), (128, 64, 1000, 2, -8), (256, 256, 1000, 2, 0), (256, 64, 777, 3, 0),
                                      (128, 128, 4200, 1, -512), (256, 256, 16000, 4, 0)]:
        dY = torch.randn(nseq, rows, M_, device="cuda", generator=g)
        X = torch.randn(nseq, rows, Cx, device="cuda", generator=g)
        dYs, Xs = split(dY), split(X)
        dW = torch.zeros(M_, Cx, device="cuda")
        rc = lib.wn_tcs_debug_wgrad(h, ptr(dYs), M_, ptr(Xs), Cx, rows, nseq, off, 0.5, ptr(dW), st)
        assert rc == 0, lib.wn_last_error()
        torch.cuda.synchronize()
        X64 = unsplit(Xs).double()
        Xsh = torch.zeros_like(X64)
        if off == 0:
            Xsh = X64
        else:
            Xsh[:, -off:] = X64[:, :off]
        ref = 0.5 * torch.einsum("brm,brc->mc", unsplit(dYs).double(), Xsh)
        err = ((dW.double() - ref).abs().max() / ref.abs().max()).item()
        rowerr = (dW.double() - ref).abs().max(dim=1).values / ref.abs().max()
        worst = max(worst, err)
        print("wgrad M=%d C=%d rows=%d seq=%d off=%d: max rel-to-max err %.2e  (worst rows %s)" %
              (M_, Cx, rows, nseq, off, err, torch.topk(rowerr, 4).indices.tolist()))
    print("WORST", worst)


if __name__ == "__main__":
    main()
