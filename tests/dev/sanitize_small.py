"""Small end-to-end exercise for compute-sanitizer (manual): fp32 + tf32 train step, generator v1/v2/v3."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net
for name, B, W in (("tiny_k3_bias", 2, 45), ("C_small", 2, 300)):
    cfg = make_cfg(name)
    w = O.init_weights(cfg, np.random.default_rng(0), np.float64, bias_scale=0.1 if name != "C_small" else 0.0)
    Q = cfg.quantization_steps
    x = torch.from_numpy(np.random.default_rng(1).integers(0, Q, (B, W + 1)).astype(np.int32)).cuda()
    for prec in ("fp32", "tf32"):
        net = make_net(cfg, w); net.set_precision(prec)
        loss = net.train_step(x[:, :W].contiguous(), x[:, 1:].contiguous(), train_width=W - 7)
        torch.cuda.synchronize()
        print(name, prec, float(loss[0]), flush=True)
    gnet = make_net(cfg, w, faster=True)
    win = np.random.default_rng(2).integers(0, Q, (3, O.input_width(cfg))).astype(np.int32)
    out = gnet.generate(win, 12, mode="sample", seed=3)
    torch.cuda.synchronize()
    print(name, "gen", out[0, :6].tolist(), flush=True)
cfg = make_cfg("C")
w = O.init_weights(cfg, np.random.default_rng(0), np.float32)
# full-depth config C, ragged width, T < W: fused layer kernel (TMA stores), fused gate-backward, grouped dzs/dWs
x = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (3, 1301)).astype(np.int32)).cuda()
net = make_net(cfg, w); net.set_precision("tf32")
for _ in range(2):
    loss = net.train_step(x[:, :1300].contiguous(), x[:, 1:].contiguous()[:, -900:].contiguous(), train_width=900)
torch.cuda.synchronize()
print("C tf32 train", float(loss[0]), flush=True)
gnet = make_net(cfg, w, faster=True)
win = np.random.default_rng(2).integers(0, 256, (2, O.input_width(cfg))).astype(np.int32)
print("C gen v3", gnet.generate(win, 6, mode="greedy")[0].tolist(), flush=True)
