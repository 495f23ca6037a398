"""Manual check: numerical error of the tensor-core BACKWARD alone (same TF32 forward tape, TC backward vs exact-fp32
SIMT backward), and gradient error vs the oracle when the targets are learnable (structured) instead of random."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net, rel_err
from wavenet_b200 import _lib
cfg = make_cfg("C_small")
w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
B, W = 2, 1000
x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
tgt = np.random.default_rng(1).integers(0, 256, (B, W)).astype(np.int32)
net = make_net(cfg, w); net.set_precision("tf32")
logits = net.forward_one_step(x, apply_softmax=False); loss = net.cross_entropy(logits, tgt)
net.backward(); g_tc = net.get_grads()
# same tape, exact-fp32 SIMT backward (it recomputes the gates from x in fp32)
_lib.check(net._libh.wn_set_precision(net._h, _lib.WN_PREC_FP32))
net.backward(); g_simt = net.get_grads()
errs = sorted(((rel_err(g_tc[k], g_simt[k]), k) for k in g_tc if np.abs(g_simt[k]).max() > 0), reverse=True)
print("TC backward vs fp32 SIMT backward on the SAME TF32 tape: worst %.3e (%s), median %.3e" % (errs[0][0], errs[0][1], np.median([e[0] for e in errs])))
fw = O.forward_loss(cfg, w, x, tgt, dtype=np.float64); g_ref = O.backward(cfg, fw)
errs2 = sorted(((rel_err(g_simt[k], g_ref[k]), k) for k in g_ref if np.abs(g_ref[k]).max() > 0), reverse=True)
print("fp32 SIMT backward on the TF32 tape vs fp64 oracle (forward error only): worst %.3e, median %.3e" % (errs2[0][0], np.median([e[0] for e in errs2])))
