import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from wavenet_b200 import _lib as _L
_L.SIGNATURES.pop("wn_accumulate_grads", None)          # older builds (bisecting with WN_LIB_PATH)
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net, rel_err
cfg = make_cfg("C_small")
for seed in (0, 5, 11):
    w = O.init_weights(cfg, np.random.default_rng(seed), np.float64)
    x = np.random.default_rng(1).integers(0, 256, (2, 300)).astype(np.int32)
    tgt = np.random.default_rng(2).integers(0, 256, (2, 300)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, tgt, dtype=np.float64)
    g_ref = O.backward(cfg, fw)
    net = make_net(cfg, w); net.set_precision("fp16x2")
    logits = net.forward_one_step(x, apply_softmax=False)
    loss = net.cross_entropy(logits, tgt)
    net.backward()
    g = net.get_grads()
    errs = sorted(((rel_err(g[k], v), k, float(np.linalg.norm(v))) for k, v in g_ref.items() if np.abs(v).max() > 0), reverse=True)
    print("seed", seed, [(("%.2e" % e), k, "%.2e" % n) for e, k, n in errs[:6]])
