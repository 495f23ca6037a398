"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE.

oracle/ref_build.py turns /root/reference/{wavenet,faster_wavenet,data}.py into Python 3 (mechanical transform, listed
there) and oracle/chainer_shim/ supplies the Chainer primitives in NumPy; everything else that runs below is the
reference: DilatedConvolution1D.__call__ with its pad / reshape / slice (wavenet.py:294-342), ResidualConvLayer
(:358-368), forward_causal/residual/softmax_block (:565-593), slice_1d (:531), cross_entropy (:597-617), backprop with
GradientClipping / WeightDecay hooks and Adam (:175-199, 457-519), FasterWaveNet._forward_one_step (faster_wavenet.py:
50-113, incl. the ELU head and rolled windows), data.load_audio_file / save_audio_file / onehot_pixel_image.

The fixtures pin oracle/wavenet_oracle.py (tests/test_reference_pin.py, CPU) and the CUDA path (tests/test_gpu_reference.py).
The reference computes in float32 (CausalPadding1d hard-codes it, wavenet.py:224), so do these fixtures.
Only runs where /root/reference exists:  python tests/golden/make_ref_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wavenet_oracle as O  # noqa: E402  (only for seeded weights and the shared configurations)
from oracle.ref_build import import_reference  # noqa: E402
from tests.util import make_cfg  # noqa: E402

from tests.util import digest, weights_from_seed  # noqa: E402

HERE = os.environ.get("WN_REF_GOLDEN_OUT") or os.path.dirname(os.path.abspath(__file__))
R, RF, RD = import_reference()

CFG_KEYS = ["quantization_steps", "causal_conv_no_bias", "causal_conv_filter_width", "causal_conv_channels",
            "residual_conv_dilation_no_bias", "residual_conv_projection_no_bias", "residual_conv_filter_width",
            "residual_conv_channels", "residual_num_blocks", "softmax_conv_no_bias", "softmax_conv_channels",
            "weight_decay", "momentum", "gradient_clipping"]


def ref_params(cfg):
    p = R.Params()
    for k in CFG_KEYS:
        setattr(p, k, getattr(cfg, k))
    return p


def inject(net, w):
    """Overwrite the randomly initialised links with the given weights (reference link names, wavenet.py:461-472)."""
    seen = set()
    for path, param in net.chain.namedparams():          # "/causal_0/W"
        name = path[1:]
        assert param.data.shape == w[name].shape, (name, param.data.shape, w[name].shape)
        param.data[...] = w[name].astype(np.float32)
        seen.add(name)
    assert seen == set(w.keys()), (sorted(set(w.keys()) ^ seen))


def forward_train(net, cfg, x, tgt, T):
    """train_audio/train.py:62-78."""
    W = x.shape[1]
    onehot = RD.onehot_pixel_image(x, quantization_steps=cfg.quantization_steps)
    causal = net.forward_causal_block(onehot)
    out, skip = net.forward_residual_block(causal)
    skip_s = net.slice_1d(skip, W - T) if W - T >= 1 else skip
    logits = net.forward_softmax_block(skip_s, apply_softmax=False)
    loss = net.cross_entropy(logits, tgt)
    return causal, out, skip, logits, loss


def train_case(tag, name, B, W, T, seed, bias_scale=0.2, full=True):
    """full=True stores every tensor; full=False (the wide networks) stores digests of the gradients / updated weights
    (norm, three seeded random projections, the first 64 values: tests/util.py:digest) and every 5th time position of
    the block outputs, so that the fixtures stay small.  Weights are never stored: tests rebuild them from the seed."""
    cfg = make_cfg(name)
    w = weights_from_seed(cfg, seed, bias_scale)
    rng = np.random.default_rng(seed + 1000)
    Q = cfg.quantization_steps
    x = rng.integers(0, Q, (B, W)).astype(np.int32)
    tgt = rng.integers(0, Q, (B, T)).astype(np.int32)
    net = R.WaveNet(ref_params(cfg))
    inject(net, w)
    net.update_laerning_rate(1e-3)                         # train_audio/train.py:95
    causal, out, skip, logits, loss = forward_train(net, cfg, x, tgt, T)
    tpos = np.arange(W) if full else np.unique(np.concatenate([np.arange(0, W, 5), np.arange(min(W, 40)), np.arange(W - 8, W)]))
    res = dict(x=x, target=tgt, T=np.int32(T), seed=np.int32(seed), bias_scale=np.float64(bias_scale), full=np.int32(full),
               tpos=tpos.astype(np.int32), causal=causal.data[:, :, 0, :][:, :, tpos], out=out.data[:, :, 0, :][:, :, tpos],
               sum_skip=skip.data[:, :, 0, :][:, :, tpos], logits=logits.data[:, :, 0, :], loss=np.float64(loss.data))
    # gradients of the loss before any hook (what loss.backward() leaves in param.grad)
    net.chain.cleargrads()
    loss.backward()
    for path, param in net.chain.namedparams():
        g = np.zeros_like(param.data) if param.grad is None else param.grad.copy()
        res["g:" + path[1:]] = g if full else digest(path[1:], g)
    # one optimiser step through the reference's backprop (hooks + Adam) on a fresh graph
    causal, out, skip, logits, loss = forward_train(net, cfg, x, tgt, T)
    net.backprop(loss)
    for path, param in net.chain.namedparams():
        res["u:" + path[1:]] = param.data.copy() if full else digest(path[1:], param.data - w[path[1:]])   # the UPDATE
    np.savez_compressed(os.path.join(HERE, "ref_train_%s.npz" % tag), **res)
    print("ref_train_%s: loss %.6f" % (tag, float(res["loss"])))


def gen_case(tag, name, steps, seed, bias_scale=0.2):
    """train_audio/generate.py:13-43 with np.argmax instead of np.random.choice (as _tests_/faster_generation/generate.py:37),
    once through _forward_one_step (fast path: ReLU on the priming call, ELU afterwards) and once through forward_one_step."""
    cfg = make_cfg(name)
    w = weights_from_seed(cfg, seed, bias_scale)
    rng = np.random.default_rng(seed + 1000)
    Q = cfg.quantization_steps
    input_width = O.input_width(cfg)
    res = dict(seed=np.int32(seed), bias_scale=np.float64(bias_scale))
    start = rng.integers(0, Q, (input_width,)).astype(np.int32)
    for mode in ("fast", "slow"):
        net = RF.FasterWaveNet(ref_params(cfg))
        inject(net, w)
        audio = start.copy()
        probs = []
        for _ in range(steps):
            window = audio[-input_width:].reshape((1, -1))
            onehot = RD.onehot_pixel_image(window, quantization_steps=Q)
            if mode == "fast":
                softmax = net._forward_one_step(onehot, apply_softmax=True, as_numpy=True)
            else:
                softmax = net.forward_one_step(onehot, apply_softmax=True, as_numpy=True)
            softmax = softmax[0, :, 0, -1]
            probs.append(softmax.copy())
            audio = np.append(audio, [np.argmax(softmax)], axis=0)
        res["probs_" + mode] = np.stack(probs)
        res["samples_" + mode] = audio[input_width:].astype(np.int32)
    res["window"] = start.reshape(1, -1)
    np.savez_compressed(os.path.join(HERE, "ref_gen_%s.npz" % tag), **res)
    print("ref_gen_%s: fast %s... slow %s..." % (tag, res["samples_fast"][:8], res["samples_slow"][:8]))


def conv_case():
    """_tests_/dilated_conv/test_conv.py:8-26 idea on the CURRENT DilatedConvolution1D signature: known input
    mod(arange, 5), all-ones weights, several (k, d, W) incl. widths where the zero prefix differs from the dilation."""
    res = {}
    cases = [(2, 1, 10), (2, 2, 10), (2, 4, 10), (2, 8, 5), (3, 3, 10), (3, 9, 33), (2, 512, 1000), (2, 256, 700), (4, 4, 10)]
    for i, (k, d, W) in enumerate(cases):
        C_in, C_out = 4, 3
        ksize = (1, k) if d == 1 else (k, 1)
        layer = R.DilatedConvolution1D(C_in, C_out, ksize, filter_width=k, dilation=d, nobias=True)
        rng = np.random.default_rng(100 + i)
        layer.W.data[...] = rng.standard_normal(layer.W.data.shape).astype(np.float32)
        x = rng.standard_normal((2, C_in, 1, W)).astype(np.float32)
        y = layer(R.Variable(x))
        res["case%d" % i] = np.array([k, d, W], dtype=np.int32)
        res["x%d" % i], res["W%d" % i], res["y%d" % i] = x, layer.W.data.copy(), y.data.copy()
        # gradient through the reference's own CausalPadding1d / CausalSlice1d backward
        gy = rng.standard_normal(y.data.shape).astype(np.float32)
        xv = R.Variable(x)
        yv = layer(xv)
        yv.grad = gy
        layer.W.cleargrad()
        yv.backward()
        res["gy%d" % i], res["gx%d" % i], res["gW%d" % i] = gy, xv.grad.copy(), layer.W.grad.copy()
    np.savez_compressed(os.path.join(HERE, "ref_dilated_conv.npz"), **res)
    print("ref_dilated_conv: %d cases" % len(cases))


def mulaw_case():
    """data.load_audio_file / save_audio_file (data.py:5-58) on synthetic WAVs: stereo int16 (cast to float by the
    reference) and mono int16 (hits the Python-2 in-place integer division, quirk Q5)."""
    from scipy.io import wavfile
    rng = np.random.default_rng(0)
    n = np.arange(6000)
    raw = (0.5 * np.sin(2 * np.pi * 440 * n / 16000) + 0.05 * rng.standard_normal(n.size)) * 32767
    raw[:300] = 0
    raw[-200:] = 0                                             # leading / trailing silence to trim
    raw = np.clip(raw, -32768, 32767).astype(np.int16)
    stereo = np.stack([raw, raw[::-1]], axis=1)
    res = dict(stereo=stereo, mono=raw)
    with tempfile.TemporaryDirectory() as d:
        wavfile.write(os.path.join(d, "s.wav"), 16000, stereo)
        wavfile.write(os.path.join(d, "m.wav"), 16000, raw)
        q_s, sr = RD.load_audio_file(os.path.join(d, "s.wav"))
        q_m, _ = RD.load_audio_file(os.path.join(d, "m.wav"))
        res["q_stereo"], res["q_mono"], res["sr"] = q_s, q_m, np.int32(sr)
        allq = np.arange(256, dtype=np.int32)
        with np.errstate(all="ignore"):
            RD.save_audio_file(os.path.join(d, "o.wav"), allq, 256, format="16bit_pcm", sampling_rate=16000)
        sr2, pcm = wavfile.read(os.path.join(d, "o.wav"))
        res["pcm_all"], res["sr_out"] = pcm, np.int32(sr2)
    x = rng.integers(0, 256, (3, 17)).astype(np.int32)
    res["onehot_x"], res["onehot"] = x, RD.onehot_pixel_image(x, 256)
    np.savez_compressed(os.path.join(HERE, "ref_mulaw.npz"), **res)
    print("ref_mulaw: stereo %d samples kept of %d, mono %d kept (classes %s)" %
          (q_s.size, raw.size, q_m.size, np.unique(q_m).tolist()))


if __name__ == "__main__" and "--only" in sys.argv:          # subset used by tests/test_reference_pin.py
    np.random.seed(0)
    train_case("tiny_k2", "tiny_k2", 3, 37, 20, 1)
    gen_case("tiny_k2", "tiny_k2", 25, 7)
    conv_case()
    mulaw_case()
elif __name__ == "__main__":
    np.random.seed(0)
    train_case("tiny_k2", "tiny_k2", 3, 37, 20, 1)
    train_case("tiny_k3_bias", "tiny_k3_bias", 2, 61, 61, 2)
    train_case("odd", "odd", 2, 50, 33, 3)
    train_case("C_small", "C_small", 2, 300, 173, 4, bias_scale=0.0, full=False)       # config-C widths (fused tensor-core shape)
    train_case("C_small_full", "C_small", 1, 257, 257, 5, bias_scale=0.0, full=False)  # T == W: zero-prefix rows are in the loss
    train_case("B", "B", 1, 300, 43, 6, bias_scale=0.0, full=False)                    # reference default model.py network
    gen_case("tiny_k2", "tiny_k2", 25, 7)
    gen_case("tiny_k3_bias", "tiny_k3_bias", 25, 8)
    gen_case("C_small", "C_small", 20, 9, bias_scale=0.0)
    conv_case()
    mulaw_case()
    print("reference-executed fixtures written to", HERE)
