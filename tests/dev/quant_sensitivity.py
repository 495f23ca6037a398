"""CPU experiment (torch fp64 autograd, no GPU): which tensors of the train step tolerate ONE fp16 plane (round to nearest)
under the 1e-3 gradient gate.  usage: python tests/dev/quant_sensitivity.py C 1 4200 1129 [variant ...].
Result (config C): forward operands (x, z, head activations) need the split format (a single plane moves the logits by 2e-4 and
the gradients by 2e-2); every backward tensor alone costs 2-5e-4; dafg + dzs + dskip + dh + dlogits together 6.4e-4, with dout
as well 1.2e-3 (fails); a fp16 sigmoid tape adds 6.6e-4 (hence the 16-bit fixed-point tape)."""
import sys, itertools
import numpy as np, torch
torch.set_num_threads(16)
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import wavenet_oracle as O
from tests.util import make_cfg

def r16(x, scale=1.0):
    return (x * scale).to(torch.float16).to(torch.float64) / scale
def r11t(x):  # truncation to 10 explicit mantissa bits (tf32 MMA behaviour)
    xf = x.to(torch.float32).contiguous()
    i = xf.view(torch.int32) & ~0x1FFF
    return i.view(torch.float32).to(torch.float64)

class QF(torch.autograd.Function):   # forward rounding, straight-through backward
    @staticmethod
    def forward(ctx, x, mode): return MODES[mode](x)
    @staticmethod
    def backward(ctx, g): return g, None
class QB(torch.autograd.Function):   # identity forward, gradient rounded in backward
    @staticmethod
    def forward(ctx, x, mode, scale): ctx.mode, ctx.scale = mode, scale; return x.clone()
    @staticmethod
    def backward(ctx, g): return MODES[ctx.mode](g * ctx.scale) / ctx.scale, None, None
class Gate(torch.autograd.Function):   # z = tanh(af) sigmoid(ag); backward from the tape (z exact, sigmoid rounded), tanh = z / sigmoid
    @staticmethod
    def forward(ctx, af, ag, mode):
        th, sg = torch.tanh(af), torch.sigmoid(ag)
        z = th * sg
        ctx.save_for_backward(z, MODES[mode](sg))
        return z
    @staticmethod
    def backward(ctx, dz):
        z, sg = ctx.saved_tensors
        th = z / sg
        return dz * sg * (1 - th * th), dz * th * sg * (1 - sg), None
MODES = {"none": lambda x: x, "f16": r16, "trunc": r11t, "f32": lambda x: x.to(torch.float32).to(torch.float64)}

def run(cfg, w, x, tgt, T, q, gscale=1.0):
    """q: dict tensor-name -> mode"""
    g = lambda k: q.get(k, "none")
    P = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in w.items()}
    B, W = x.shape
    xi = torch.tensor(x, dtype=torch.int64)
    Wc = P["causal_0/W"].reshape(64, 256, 2)
    # causal: out[t] = W[:, x[t-1], 0] + W[:, x[t], 1]
    e1 = Wc[:, :, 1].t()[xi]                     # B,W,C
    e0 = Wc[:, :, 0].t()[xi]
    e0 = torch.cat([torch.zeros_like(e0[:, :1]), e0[:, :-1]], 1)
    h = (e0 + e1)                                # B,W,64
    layers = O._layers(cfg)
    skip = 0
    zs = []
    for li, (name, d) in enumerate(LAYERS):
        hq = QF.apply(h, g("x_op"))            # operand rounding of the residual stream into the gate GEMM
        hq = QB.apply(hq, g("dx_part"), gscale)
        hs = torch.cat([torch.zeros_like(hq[:, :d]), hq[:, :-d]], 1)
        Wf, Wg = (P[name + s].reshape(64, 64, 2) for s in ("_wf/W", "_wg/W")); Wp = P[name + "_projection_block/W"].reshape(64, 64, 1); Ws = P[name + "_projection_softmax/W"].reshape(256, 64, 1)
        af = hs @ Wf[:, :, 0].t() + hq @ Wf[:, :, 1].t()
        ag = hs @ Wg[:, :, 0].t() + hq @ Wg[:, :, 1].t()
        af = QB.apply(af, g("dafg"), gscale); ag = QB.apply(ag, g("dafg"), gscale)
        th = torch.tanh(af); sg = torch.sigmoid(ag)
        z = Gate.apply(af, ag, g("sg"))
        z = QF.apply(z, g("z"))
        zsk = QB.apply(z, g("dzs"), gscale)
        skip = skip + zsk @ Ws[:, :, 0].t()
        zp = QB.apply(z, g("dz_p"), gscale)
        h = h + zp @ Wp[:, :, 0].t()
        h = QF.apply(h, g("x_store"))
        h = QB.apply(h, g("dout"), gscale)
    s = skip[:, W - T:]
    s = QB.apply(s, g("dskip"), gscale)
    a0 = QF.apply(torch.relu(s), g("h0"))
    h1 = a0 @ P["softmax_0/W"][:, :, 0, 0].t()
    h1 = QB.apply(h1, g("dh1"), gscale)
    a1 = QF.apply(torch.relu(h1), g("h1"))
    lg = a1 @ P["softmax_1/W"][:, :, 0, 0].t()
    lg = QB.apply(lg, g("dlogits"), gscale)
    loss = torch.nn.functional.cross_entropy(lg.reshape(-1, 256), torch.tensor(tgt, dtype=torch.int64).reshape(-1))
    loss.backward()
    return lg.detach().numpy(), {k: v.grad.numpy() for k, v in P.items() if v.grad is not None}

def rel(a, b): return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)

if __name__ == "__main__":
    name, B, W, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    cfg = make_cfg(name)
    LAYERS = []
    for b in range(cfg.residual_num_blocks):
        for l in range(len(cfg.residual_conv_channels)):
            LAYERS.append(("residual_%d_block_%d" % (b, l), 2 ** l))
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
    print([k for k in w][:8])
    x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
    tgt = np.random.default_rng(1).integers(0, 256, (B, T)).astype(np.int32)
    lg0, g0 = run(cfg, w, x, tgt, T, {})
    gs = float(B * T)
    variants = {
        "z": {"z": "f16"}, "sg": {"sg": "f16"}, "x_op": {"x_op": "f16"}, "x_store": {"x_store": "f16"},
        "x_op_trunc": {"x_op": "trunc"},
        "dafg": {"dafg": "f16"}, "dzs": {"dzs": "f16"}, "dout": {"dout": "f16"}, "dz_p": {"dz_p": "f16"}, "dx_part": {"dx_part": "f16"},
        "dafg_trunc": {"dafg": "trunc"}, "dout_trunc": {"dout": "trunc"},
        "h0": {"h0": "f16"}, "h1": {"h1": "f16"}, "dskip": {"dskip": "f16"}, "dh1": {"dh1": "f16"}, "dlogits": {"dlogits": "f16"},
        "tapes(z,sg,dafg,dzs)": {"z": "f16", "sg": "f16", "dafg": "f16", "dzs": "f16"},
        "bwd_no_dout": {"dafg": "f16", "dzs": "f16", "dskip": "f16", "dh1": "f16", "dlogits": "f16"},
        "bwd_no_dout+sg": {"sg": "f16", "dafg": "f16", "dzs": "f16", "dskip": "f16", "dh1": "f16", "dlogits": "f16"},
        "bwd_all+sg": {"sg": "f16", "dafg": "f16", "dzs": "f16", "dout": "f16", "dskip": "f16", "dh1": "f16", "dlogits": "f16"},
        "bwd_all": {"dafg": "f16", "dzs": "f16", "dout": "f16", "dskip": "f16", "dh1": "f16", "dlogits": "f16"},
        "all_f16": {k: "f16" for k in ("z", "sg", "x_store", "dafg", "dzs", "dout", "h0", "h1", "dskip", "dh1", "dlogits")},
    }
    sel = sys.argv[5:] or list(variants)
    for vn in sel:
        lg, g = run(cfg, w, x, tgt, T, variants[vn], gs)
        errs = {k: rel(g[k], g0[k]) for k in g0}
        worst = max(errs, key=errs.get)
        print("%-24s logit maxabs %.2e  grad rel: max %.2e (%s) median %.2e" % (vn, np.abs(lg - lg0).max(), errs[worst], worst, np.median(list(errs.values()))), flush=True)
