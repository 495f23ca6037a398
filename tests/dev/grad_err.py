"""Prints the fp16x2 path's logit / gradient errors against the fp64 oracle (the numbers behind tests/test_gpu_fp16x2.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net, rel_err
from tests.test_gpu_fp16x2 import run_train_step

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x2"
for name, B, W, T in (("C_small", 2, 1000, 1000), ("C", 1, 4200, 1129), ("C", 3, 2171, 977), ("B", 1, 600, 343)):
    cfg = make_cfg(name)
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
    x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
    tgt = np.random.default_rng(1).integers(0, 256, (B, T)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, tgt, train_width=T, dtype=np.float64)
    g_ref = O.backward(cfg, fw)
    net = make_net(cfg, w)
    net.set_precision(prec)
    logits, loss = run_train_step(net, x, tgt, T)
    got = logits.data.detach().cpu().numpy()[:, :, 0, :]
    net.backward()
    g = net.get_grads()
    errs = {k: rel_err(g[k], v) for k, v in g_ref.items() if np.abs(v).max() > 0}
    worst = max(errs, key=errs.get)
    print("%s %dx%d T=%d [%s]: logits %.2e  loss %.2e  grad max %.2e (%s) median %.2e" % (
        name, B, W, T, prec, np.abs(got - fw["logits"]).max(), abs(float(loss.data) - float(fw["loss"])), errs[worst], worst,
        np.median(list(errs.values()))), flush=True)
    # the fused train-step path (wn_forward_loss: fused CE epilogue, single-plane head gradients)
    net = make_net(cfg, w)
    net.set_precision(prec)
    net._bind(B, W)
    net._fwd_bwd(torch.from_numpy(x).cuda(), torch.from_numpy(tgt).cuda(), T)
    g = net.get_grads()
    errs = {k: rel_err(g[k], v) for k, v in g_ref.items() if np.abs(v).max() > 0}
    worst = max(errs, key=errs.get)
    print("    train_step path: loss %.2e  grad max %.2e (%s) median %.2e" % (
        abs(float(net._loss[0]) - float(fw["loss"])), errs[worst], worst, np.median(list(errs.values()))), flush=True)
