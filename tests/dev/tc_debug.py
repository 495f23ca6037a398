"""Manual diagnostic (not a pytest): TF32 tensor-core path vs the exact-fp32 SIMT path, stage by stage."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net

name = sys.argv[1] if len(sys.argv) > 1 else "C_small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
cfg = make_cfg(name)
w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
res = {}
for prec in ("fp32", "tf32"):
    net = make_net(cfg, w)
    net.set_precision(prec)
    print(prec, "tc_active =", net._libh.wn_tc_active(net._h), flush=True)
    c = net.forward_causal_block(x)
    out, skip = net.forward_residual_block(c)
    torch.cuda.synchronize()
    print(prec, "residual done", flush=True)
    logits = net.forward_softmax_block(skip, apply_softmax=False)
    torch.cuda.synchronize()
    res[prec] = (out.data.cpu().numpy(), skip.data.cpu().numpy(), logits.data.cpu().numpy())
for i, nm in enumerate(("out", "skip", "logits")):
    a, b = res["fp32"][i], res["tf32"][i]
    err = np.abs(a - b)
    print("%-7s max|fp32| %.3e  max err %.3e  mean err %.3e  worst t=%d" % (nm, np.abs(a).max(), err.max(), err.mean(),
          int(np.unravel_index(err.argmax(), err.shape)[3])))
    # error profile along time (first 8 blocks of 128)
    prof = [err[..., k * 128:(k + 1) * 128].max() for k in range(min(8, (W + 127) // 128))]
    print("        per-128 max err:", " ".join("%.1e" % p for p in prof))

# ---- gradients: tensor-core backward vs fp32 SIMT backward vs fp64 oracle ----
T = W
tgt = np.random.default_rng(1).integers(0, 256, (B, T)).astype(np.int32)
fw = O.forward_loss(cfg, w, x, tgt, train_width=T, dtype=np.float64)
g_ref = O.backward(cfg, fw)
from tests.util import rel_err
grads = {}
for prec in ("fp32", "tf32"):
    net = make_net(cfg, w)
    net.set_precision(prec)
    logits = net.forward_one_step(x, apply_softmax=False)
    loss = net.cross_entropy(logits, tgt)
    net.backward()
    torch.cuda.synchronize()
    grads[prec] = net.get_grads()
    print(prec, "loss", float(loss.data), "oracle", float(fw["loss"]))
worst = []
for k, v in g_ref.items():
    if np.abs(v).max() == 0:
        worst.append((0.0, k, float(np.abs(grads["tf32"][k]).max())))
        continue
    worst.append((rel_err(grads["tf32"][k], v), k, rel_err(grads["fp32"][k], v)))
worst.sort(reverse=True)
print("worst tf32 grad rel errs (tf32, name, fp32):")
for e in worst[:12]:
    print("  %.3e  %-44s %.3e" % e)
print("median tf32 rel err %.3e" % np.median([e[0] for e in worst]))
for k in ["softmax_1/W", "softmax_1/b", "softmax_0/W", "residual_1_block_5_wf/W", "residual_1_block_5_projection_softmax/W", "residual_0_block_0_projection_block/W", "causal_0/W"]:
    a, b_, r = grads["tf32"][k], grads["fp32"][k], g_ref[k]
    print("%-44s |tf32| %.4e |fp32| %.4e |ref| %.4e  tf32[0,:4]=%s ref[0,:4]=%s" % (k, np.linalg.norm(a), np.linalg.norm(b_), np.linalg.norm(r), a.reshape(a.shape[0], -1)[0, :4] if a.ndim > 1 else a[:4], r.reshape(r.shape[0], -1)[0, :4] if r.ndim > 1 else r[:4]))
