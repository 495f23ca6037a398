"""Shared helpers for the parity tests."""
import numpy as np

from oracle import wavenet_oracle as O


def make_cfg(name):
    if name == "tiny_k2":
        return O.OracleParams(quantization_steps=6, causal_conv_channels=[4], residual_conv_channels=[3, 3, 3],
                              residual_num_blocks=2, softmax_conv_channels=[5, 6])
    if name == "tiny_k3_bias":
        return O.OracleParams(quantization_steps=6, causal_conv_channels=[4, 5], causal_conv_filter_width=3,
                              causal_conv_no_bias=False, residual_conv_filter_width=3,
                              residual_conv_dilation_no_bias=False, residual_conv_projection_no_bias=False,
                              residual_conv_channels=[3, 2, 3], residual_num_blocks=2,
                              softmax_conv_channels=[5, 7, 6])
    if name == "odd":   # channel counts that are not multiples of anything
        return O.OracleParams(quantization_steps=37, causal_conv_channels=[19], residual_conv_channels=[13, 11, 13, 7],
                              residual_num_blocks=2, softmax_conv_channels=[23, 37], weight_decay=0.01)
    if name == "A":     # Params() defaults, wavenet.py:102-146
        return O.OracleParams()
    if name == "B":     # train_audio/model.py:24-43
        return O.config_B()
    if name == "C":     # BASELINE.json configs 2-5
        return O.config_C()
    if name == "C_small":  # config C channel widths, 2 blocks of 6 layers (d=1..32)
        return O.OracleParams(causal_conv_channels=[64], residual_conv_channels=[64] * 6, residual_num_blocks=2,
                              softmax_conv_channels=[256, 256, 256])
    raise KeyError(name)


def to_product_params(cfg):
    from wavenet_b200.wavenet import Params
    p = Params()
    for k in p.to_dict():
        setattr(p, k, getattr(cfg, k))
    return p


def make_net(cfg, w, faster=False, **kw):
    from wavenet_b200.wavenet import WaveNet
    from wavenet_b200.faster_wavenet import FasterWaveNet
    net = (FasterWaveNet if faster else WaveNet)(to_product_params(cfg), seed=0, **kw)
    net.set_weights(w)
    net.to_gpu()
    return net


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def weights_from_seed(cfg, seed, bias_scale=0.0):
    """float32 weights the reference-executed fixtures (tests/golden/make_ref_golden.py) were generated with."""
    w = O.init_weights(cfg, np.random.default_rng(seed), np.float64, bias_scale=bias_scale)
    return {k: v.astype(np.float32) for k, v in w.items()}


NPROJ = 16


def digest(name, a, head=64):
    """Compact fingerprint of a tensor: [L2 norm, NPROJ seeded random projections, first `head` values] in float64.
    The projections use unit-variance vectors drawn from a generator seeded by the tensor's NAME, so any party can
    recompute them; compare with digest_err()."""
    import zlib
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    proj = [float(a @ rng.standard_normal(a.size)) for _ in range(NPROJ)]
    return np.concatenate([[np.linalg.norm(a)], proj, a[:head]])


def digest_err(got, want):
    """Relative L2 error estimated from a digest.  For an error tensor e every projection difference is ~N(0, |e|^2), so the
    RMS over the NPROJ projections estimates |e| (+-18 % at 16 projections) -- the same quantity rel_err() measures on full
    tensors; the norm and the leading values are checked against the tensor's norm as well."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    scale = max(abs(float(want[0])), 1e-30)
    d = got - want
    e_proj = float(np.sqrt(np.mean(d[1:1 + NPROJ] ** 2)))
    e_rest = float(np.abs(np.concatenate([d[:1], d[1 + NPROJ:]])).max())
    return max(e_proj, e_rest) / scale


def grad_errors(g, g_ref):
    """Per-tensor relative L2 errors (tensors whose reference gradient is exactly zero must be exactly zero: error inf)."""
    out = {}
    for k, v in g_ref.items():
        if np.abs(v).max() == 0:
            out[k] = 0.0 if np.abs(g[k]).max() == 0 else float("inf")
        else:
            out[k] = rel_err(g[k], v)
    return out


def grads_vs_oracle(cfg, fw, g, eps=3e-6, max_kinks=4):
    """Worst per-tensor gradient error against the fp64 oracle, ReLU-kink aware.

    The head's ReLUs (wavenet.py:588) are not differentiable at 0: a pre-activation within the forward tolerance of zero
    (|v| < eps; with ~5e5 ReLU inputs per case the smallest one is typically ~5e-7) may legitimately fall on the other side in
    fp32 arithmetic -- the Chainer fp32 reference does the same against an fp64 run -- and ONE flipped mask moves every
    upstream gradient by ~1/sqrt(#active elements) ~ 2e-3.  The product is therefore compared with the oracle's gradient for
    the best assignment of signs to those (at most max_kinks) near-zero pre-activations; everything else is untouched.
    Returns (worst error, number of flipped kinks, dict of per-tensor errors)."""
    import itertools
    errs = grad_errors(g, O.backward(cfg, fw))
    best = (max(errs.values()), 0, errs)
    kinks = [(i, int(j)) for i, h in enumerate(fw["h_cache"]) for j in np.flatnonzero(np.abs(h) < eps)]
    if not kinks or len(kinks) > max_kinks:
        return best
    for r in range(1, len(kinks) + 1):
        for subset in itertools.combinations(kinks, r):
            fw2 = dict(fw)
            fw2["h_cache"] = [h.copy() for h in fw["h_cache"]]
            for i, j in subset:
                fw2["h_cache"][i].flat[j] = -fw2["h_cache"][i].flat[j]
            e2 = grad_errors(g, O.backward(cfg, fw2))
            if max(e2.values()) < best[0]:
                best = (max(e2.values()), r, e2)
    return best
