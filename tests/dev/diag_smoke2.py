import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from wavenet_b200 import _lib as _L
_L.SIGNATURES.pop("wn_accumulate_grads", None)
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net, rel_err
cfg = make_cfg("C_small")
w = O.init_weights(cfg, np.random.default_rng(0), np.float64)
x = np.random.default_rng(1).integers(0, 256, (2, 300)).astype(np.int32)
tgt = np.random.default_rng(2).integers(0, 256, (2, 300)).astype(np.int32)
fw = O.forward_loss(cfg, w, x, tgt, dtype=np.float64)
g_ref = O.backward(cfg, fw)
net = make_net(cfg, w); net.set_precision("fp16x2")
c = net.forward_causal_block(x)
out, skip = net.forward_residual_block(c)
print("out err %.2e skip err %.2e" % (np.abs(out.data.cpu().numpy()[:, :, 0, :] - fw["out"]).max(), np.abs(skip.data.cpu().numpy()[:, :, 0, :] - fw["sum_skip"]).max()))
logits = net.forward_softmax_block(skip, apply_softmax=False)
print("logits err %.2e" % np.abs(logits.data.cpu().numpy()[:, :, 0, :] - fw["logits"]).max())
loss = net.cross_entropy(logits, tgt)
net.backward()
g = net.get_grads()
for k, v in g_ref.items():
    if np.abs(v).max() > 0 and ("block_0_" in k or "block_5" in k or "softmax" in k or "causal" in k):
        print("%-45s %.2e" % (k, rel_err(g[k], v)))
# where do the gate pre-activations live?  (saturation statistics of the oracle)
import itertools
zs = fw.get("z") if isinstance(fw, dict) else None
print([k for k in fw.keys()][:30])
