"""Generates tests/golden/*.npz from the CPU oracle (fp64).

The reference itself cannot run in this image (Python 2 + Chainer 2) and
ships no golden vectors, so these fixtures pin the ORACLE (regression) and give the
GPU tests size-small known answers; they are not outputs of the reference.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wavenet_oracle as O  # noqa: E402
from oracle import data_oracle as D  # noqa: E402
from tests.util import make_cfg  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def train_case(name, B, W, T, seed):
    cfg = make_cfg(name)
    rng = np.random.default_rng(seed)
    w = O.init_weights(cfg, rng, np.float64, bias_scale=0.2)
    Q = cfg.quantization_steps
    x = rng.integers(0, Q, (B, W)).astype(np.int32)
    tgt = rng.integers(0, Q, (B, T)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, tgt, train_width=T, dtype=np.float64)
    g = O.backward(cfg, fw)
    w2 = {k: v.copy() for k, v in w.items()}
    st = O.new_adam_state(w2)
    norm = O.clip_and_adam(cfg, w2, {k: v.copy() for k, v in g.items()}, st, lr=1e-3)
    out = dict(x=x, target=tgt, logits=fw["logits"], loss=np.float64(fw["loss"]), norm=np.float64(norm))
    for k, v in w.items():
        out["w:" + k] = v
        out["g:" + k] = g[k]
        out["u:" + k] = w2[k]
    np.savez_compressed(os.path.join(HERE, "train_%s.npz" % name), **out)


def gen_case(name, n, steps, seed):
    cfg = make_cfg(name)
    rng = np.random.default_rng(seed)
    w = O.init_weights(cfg, rng, np.float64, bias_scale=0.2)
    Win = O.input_width(cfg)
    window = rng.integers(0, cfg.quantization_steps, (n, Win)).astype(np.int32)
    out = dict(window=window)
    for act in ("reference", "relu"):
        out["greedy_" + act] = O.RingGenerator(cfg, w, n, head_act=act, dtype=np.float64).generate_greedy(window, steps)
    for k, v in w.items():
        out["w:" + k] = v
    np.savez_compressed(os.path.join(HERE, "gen_%s.npz" % name), **out)


def mulaw_case():
    rng = np.random.default_rng(0)
    n = np.arange(4000)
    raw = (0.5 * np.sin(2 * np.pi * 440 * n / 16000) + 0.05 * rng.standard_normal(n.size)) * 32767
    raw = np.clip(raw, -32768, 32767).astype(np.int16)
    stereo = np.stack([raw, raw[::-1]], axis=1)
    q = D.encode(stereo)
    pcm = D.decode(np.arange(256))
    np.savez_compressed(os.path.join(HERE, "mulaw.npz"), stereo=stereo, q=q, pcm_all=pcm,
                        q_mono=D.encode(raw))


if __name__ == "__main__":
    train_case("tiny_k2", 3, 37, 20, 1)
    train_case("tiny_k3_bias", 2, 61, 61, 2)
    train_case("odd", 2, 50, 33, 3)
    gen_case("tiny_k2", 3, 40, 4)
    gen_case("tiny_k3_bias", 2, 40, 5)
    mulaw_case()
    print("golden fixtures written to", HERE)
