"""Developer check: forward error of the fp16x2 path against the fp64 oracle, by time position."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net

for name, B, W in [("C_small", 2, 1000), ("C", 1, 4200)]:
    cfg = make_cfg(name)
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
    x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, None, dtype=np.float64)
    for prec in ("fp32", "fp16x2"):
        net = make_net(cfg, w)
        net.set_precision(prec)
        c = net.forward_causal_block(x)
        out, skip = net.forward_residual_block(c)
        lg = net.forward_softmax_block(skip, apply_softmax=False)
        for key, got in (("out", out), ("sum_skip", skip), ("logits", lg)):
            e = np.abs(got.data.cpu().numpy()[:, :, 0, :] - fw[key]).max(axis=(0, 1))   # per position
            edges = [0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, W]
            parts = ["[%d,%d) %.1e" % (a, b, e[a:min(b, W)].max()) for a, b in zip(edges[:-1], edges[1:]) if a < W]
            print(name, prec, key, "max %.1e at t=%d |" % (e.max(), int(e.argmax())), " ".join(parts), flush=True)
