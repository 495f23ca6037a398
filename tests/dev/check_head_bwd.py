"""Developer check: head backward of the fp16x2 path against a torch fp64 recomputation from the path's own skip sum."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net, rel_err


def run(name, B, W, T, prec="fp16x2"):
    cfg = make_cfg(name)
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
    x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
    tgt = np.random.default_rng(1).integers(0, 256, (B, T)).astype(np.int32)
    net = make_net(cfg, w)
    net.set_precision(prec)
    out = net.forward_causal_block(x)
    out, skip = net.forward_residual_block(out)
    sk = skip.data.clone()                                  # (B, S, 1, W) fp32
    if W - T >= 1:
        skip = net.slice_1d(skip, W - T)
    logits = net.forward_softmax_block(skip, apply_softmax=False)
    net.cross_entropy(logits, tgt)
    net.backward()
    g = net.get_grads()
    s = sk[:, :, 0, W - T:].permute(0, 2, 1).reshape(B * T, -1).double().clamp_min(0)
    W0 = torch.from_numpy(w["softmax_0/W"][:, :, 0, 0]).cuda()
    b0 = torch.from_numpy(w["softmax_0/b"]).cuda()
    W1 = torch.from_numpy(w["softmax_1/W"][:, :, 0, 0]).cuda()
    b1 = torch.from_numpy(w["softmax_1/b"]).cuda()
    h1 = (s @ W0.T + b0).clamp_min(0)
    lg = h1 @ W1.T + b1
    got_lg = logits.data[:, :, 0, :].permute(0, 2, 1).reshape(B * T, -1).double()
    p = torch.softmax(lg, dim=1)
    p[torch.arange(B * T), torch.from_numpy(tgt.reshape(-1)).cuda().long()] -= 1
    dl = p / (B * T)
    dh1 = (dl @ W1) * (h1 > 0)
    ref = {"softmax_1/W": dl.T @ h1, "softmax_1/b": dl.sum(0), "softmax_0/W": dh1.T @ s, "softmax_0/b": dh1.sum(0)}
    msg = "%s B=%d W=%d T=%d %s: logits %.1e |" % (name, B, W, T, prec, (got_lg - lg).abs().max().item())
    for k, v in ref.items():
        msg += " %s %.1e" % (k, rel_err(g[k].reshape(v.shape), v.cpu().numpy()))
    e = np.abs(g["softmax_0/W"][:, :, 0, 0] - ref["softmax_0/W"].cpu().numpy())
    msg += " | worst rows %s cols %s" % (np.argsort(e.max(1))[-3:].tolist(), np.argsort(e.max(0))[-3:].tolist())
    print(msg, flush=True)


for args in [("C_small", 2, 1000, 1000), ("C_small", 2, 1000, 999), ("C_small", 2, 1000, 600), ("C_small", 1, 1000, 1000),
             ("C_small", 1, 1024, 1024), ("C_small", 1, 128, 128), ("C_small", 1, 160, 160), ("C_small", 2, 1000, 1000, "fp32"),
             ("C", 1, 4200, 1129), ("C", 1, 4200, 4200)]:
    run(*args)
