"""Structural validation of the CPU oracle (no GPU).

The reference pins no numeric result, so the oracle is checked
against itself in independent formulations, against torch autograd, against
finite differences and against hand-derived known inputs.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import wavenet_oracle as O
from oracle import data_oracle as D


def tiny_params(**kw):
    base = dict(quantization_steps=6, causal_conv_channels=[4], residual_conv_channels=[3, 3, 3],
                residual_num_blocks=2, softmax_conv_channels=[5, 6])
    base.update(kw)
    return O.OracleParams(**base)


# ---- dilated conv: literal reshape trick == closed form (Q1) -------------------------
@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("d", [1, 2, 4, 8, 512])
@pytest.mark.parametrize("W", [5, 16, 17, 33, 257, 1000])
def test_dilated_conv_literal_equals_closed(k, d, W):
    rng = np.random.default_rng(k * 1000 + d + W)
    x = rng.standard_normal((2, 3, W))
    shape = (4, 3, 1, k) if d == 1 else (4, 3, k, 1)
    Wt = rng.standard_normal(shape)
    b = rng.standard_normal(4)
    lit = O.dilated_conv_literal(x[:, :, None, :], Wt, b, d, k)[:, :, 0, :]
    clo = O.dilated_conv_closed(x, Wt, b, d, k)
    assert lit.shape == clo.shape == (2, 4, W)
    np.testing.assert_allclose(lit, clo, rtol=0, atol=1e-12)
    zp = O.zero_prefix(W, d, k)
    assert np.all(lit[:, :, :min(zp, W)] == 0)


def test_zero_prefix_at_16000_and_3071():
    # W=16000,k=2: zp=d for d<=128, 128 for d=256 and 512.
    for d in [2, 4, 8, 16, 32, 64, 128]:
        assert O.zero_prefix(16000, d, 2) == d
    assert O.zero_prefix(16000, 256, 2) == 128
    assert O.zero_prefix(16000, 512, 2) == 128
    # input_width of config C = 3071 = 3072-1 -> pad 1 -> zp = d-1
    for d in [2, 4, 512]:
        assert O.zero_prefix(3071, d, 2) == d - 1


def test_all_ones_known_input():
    # _tests_/dilated_conv/test_conv.py:8-16: x = mod(arange(C*W),5), all-ones filter,
    # k=4; with the current code dilation = 4 -> pad 6, zp 6, last 4 outputs are plain sums.
    C, W, k, d = 4, 10, 4, 4
    x = np.mod(np.arange(0, C * W), 5).reshape((1, C, 1, W)).astype(np.float32)
    Wt = np.ones((3, C, k, 1), dtype=np.float32)
    out = O.dilated_conv_literal(x, Wt, None, d, k)
    assert out.shape == (1, 3, 1, W)
    assert O.zero_prefix(W, d, k) == 6
    assert np.all(out[0, :, 0, :6] == 0)
    for t in range(6, 10):
        want = sum(x[0, c, 0, t - (k - 1 - i) * d] for c in range(C) for i in range(k) if t - (k - 1 - i) * d >= 0)
        assert np.all(out[0, :, 0, t] == want)


# ---- pad / slice (reference _tests_/padding, _tests_/slice shapes) ---------------------
def test_padding_slice_forward_backward():
    x = np.random.default_rng(0).standard_normal((2, 2, 1, 2)).astype(np.float32)
    y = O.causal_padding_1d(x, 3)
    assert y.shape == (2, 2, 1, 5) and np.all(y[..., :3] == 0) and np.all(y[..., 3:] == x)
    gy = np.random.default_rng(1).standard_normal(y.shape).astype(np.float32)
    assert np.all(O.causal_padding_1d_backward(gy, 3) == gy[..., 3:])
    x = np.random.default_rng(2).standard_normal((2, 2, 1, 5)).astype(np.float32)
    s = O.causal_slice_1d(x, 3)
    assert np.all(s == x[..., 3:])
    gs = np.ones_like(s)
    g = O.causal_slice_1d_backward(x.shape, gs, 3)
    assert np.all(g[..., :3] == 0) and np.all(g[..., 3:] == 1)
    with pytest.raises(Exception):
        O.causal_slice_1d(x, 0)


# ---- full forward: literal == closed ----------------------------------------------------
@pytest.mark.parametrize("bias", [False, True])
def test_forward_literal_equals_closed(bias):
    p = tiny_params(causal_conv_no_bias=not bias, residual_conv_dilation_no_bias=not bias,
                    residual_conv_projection_no_bias=not bias)
    rng = np.random.default_rng(3)
    w = O.init_weights(p, rng, np.float64, bias_scale=0.3)
    x = rng.integers(0, 6, (2, 21))
    fw = O.forward_loss(p, w, x, None, dtype=np.float64)
    lit = O.forward_literal(p, w, O.onehot_pixel_image(x, 6).astype(np.float64))
    np.testing.assert_allclose(lit["y"][:, :, 0, :], fw["logits"], atol=1e-12)
    np.testing.assert_allclose(lit["out"][:, :, 0, :], fw["out"], atol=1e-12)


# ---- backward == torch autograd (independent restatement) ----------------------------------
def torch_forward_loss(p, tw, x_idx, target, train_width):
    Q = p.quantization_steps
    k, kc = p.residual_conv_filter_width, p.causal_conv_filter_width
    B, W = x_idx.shape
    h = F.one_hot(torch.as_tensor(x_idx), Q).permute(0, 2, 1).double()

    def dconv(x, Wt, b, d, kk):
        w3 = Wt.reshape(Wt.shape[0], Wt.shape[1], kk)
        y = F.conv1d(F.pad(x, ((kk - 1) * d, 0)), w3, b, dilation=d)
        zp = O.zero_prefix(W, d, kk)
        if zp > 0:
            mask = torch.ones(W, dtype=torch.float64)
            mask[:zp] = 0
            y = y * mask
        return y

    for i in range(len(p.causal_conv_channels)):
        h = dconv(h, tw["causal_{}/W".format(i)], tw.get("causal_{}/b".format(i)), 1, kc)
    skip = 0
    x = h
    for j in range(p.residual_num_blocks):
        for i in range(len(p.residual_conv_channels)):
            base = "residual_{}_block_{}_".format(j, i)
            d = k ** i
            z = torch.tanh(dconv(x, tw[base + "wf/W"], tw.get(base + "wf/b"), d, k)) * \
                torch.sigmoid(dconv(x, tw[base + "wg/W"], tw.get(base + "wg/b"), d, k))
            x = F.conv1d(z, tw[base + "projection_block/W"][:, :, 0], tw.get(base + "projection_block/b")) + x
            skip = skip + F.conv1d(z, tw[base + "projection_softmax/W"][:, :, 0], tw.get(base + "projection_softmax/b"))
    y = skip[:, :, W - train_width:]
    for i in range(len(p.softmax_conv_channels) - 1):
        y = F.conv1d(F.relu(y), tw["softmax_{}/W".format(i)][:, :, 0], tw.get("softmax_{}/b".format(i)))
    loss = F.cross_entropy(y.permute(0, 2, 1).reshape(-1, Q), torch.as_tensor(target).reshape(-1).long())
    return y, loss


@pytest.mark.parametrize("cfg", ["k2", "k3_bias_2causal"])
def test_backward_matches_torch_autograd(cfg):
    if cfg == "k2":
        p = tiny_params()
        W, T = 37, 20
    else:
        p = tiny_params(residual_conv_filter_width=3, causal_conv_channels=[4, 5], causal_conv_filter_width=3,
                        causal_conv_no_bias=False, residual_conv_dilation_no_bias=False,
                        residual_conv_projection_no_bias=False, residual_conv_channels=[3, 2, 3])
        W, T = 61, 61
    rng = np.random.default_rng(7)
    w = O.init_weights(p, rng, np.float64, bias_scale=0.2)
    x = rng.integers(0, 6, (3, W))
    tgt = rng.integers(0, 6, (3, T))
    fw = O.forward_loss(p, w, x, tgt, train_width=T, dtype=np.float64)
    g = O.backward(p, fw)
    tw = {n: torch.tensor(a, dtype=torch.float64, requires_grad=True) for n, a in w.items()}
    y, loss = torch_forward_loss(p, tw, x, tgt, T)
    loss.backward()
    np.testing.assert_allclose(fw["logits"], y.detach().numpy(), atol=1e-11)
    assert abs(float(fw["loss"]) - float(loss.detach())) < 1e-11
    for n, _ in O.param_shapes(p):
        tg = tw[n].grad
        tg = np.zeros_like(w[n]) if tg is None else tg.numpy()
        np.testing.assert_allclose(g[n], tg, atol=1e-11, err_msg=n)
    # last layer's projection_block is never reached by backward
    last = "residual_{}_block_{}_projection_block/W".format(p.residual_num_blocks - 1,
                                                            len(p.residual_conv_channels) - 1)
    assert np.all(g[last] == 0)


def test_backward_matches_finite_differences():
    p = tiny_params()
    rng = np.random.default_rng(11)
    w = O.init_weights(p, rng, np.float64)
    x = rng.integers(0, 6, (2, 19))
    tgt = rng.integers(0, 6, (2, 19))
    g = O.backward(p, O.forward_loss(p, w, x, tgt, dtype=np.float64))
    eps = 1e-6
    for name in ["causal_0/W", "residual_0_block_2_wf/W", "residual_1_block_1_wg/W",
                 "residual_0_block_0_projection_softmax/W", "softmax_0/b"]:
        flat = w[name].reshape(-1)
        for idx in rng.choice(flat.size, size=min(4, flat.size), replace=False):
            old = flat[idx]
            flat[idx] = old + eps
            lp = float(O.forward_loss(p, w, x, tgt, dtype=np.float64)["loss"])
            flat[idx] = old - eps
            lm = float(O.forward_loss(p, w, x, tgt, dtype=np.float64)["loss"])
            flat[idx] = old
            assert abs((lp - lm) / (2 * eps) - g[name].reshape(-1)[idx]) < 1e-7, name


# ---- optimiser -------------------------------------------------------------------------------
def test_clip_and_adam_known_answer():
    p = O.OracleParams(quantization_steps=2, causal_conv_channels=[1], residual_conv_channels=[1],
                       residual_num_blocks=1, softmax_conv_channels=[1, 2], gradient_clipping=1.0)
    w = {n: np.full(s, 0.5, np.float64) for n, s in O.param_shapes(p)}
    g = {n: np.full(s, 2.0, np.float64) for n, s in O.param_shapes(p)}
    n_el = sum(a.size for a in w.values())
    st = O.new_adam_state(w)
    norm = O.clip_and_adam(p, w, g, st, lr=1e-3)
    assert abs(norm - 2.0 * math.sqrt(n_el)) < 1e-12
    gc = 2.0 / norm                     # clipped gradient element
    m = 0.1 * gc
    v = 0.001 * gc * gc
    step = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    want = 0.5 - step * m / (math.sqrt(v) + 1e-8)
    for a in w.values():
        np.testing.assert_allclose(a, want, rtol=0, atol=1e-15)
    assert st["t"] == 1


# ---- generator ----------------------------------------------------------------------------------
def test_ring_generator_matches_literal_fast_generator():
    p = tiny_params(residual_conv_projection_no_bias=False)
    rng = np.random.default_rng(5)
    w = O.init_weights(p, rng, np.float64, bias_scale=0.2)
    Win = O.input_width(p)
    assert Win == (2 ** 3 - 1) * 2 + 1 + 1
    seq = list(rng.integers(0, 6, Win))
    lit = O.LiteralFastGenerator(p, w, np.float64)
    ring = O.RingGenerator(p, w, 1, head_act="reference", dtype=np.float64)
    for step in range(12):
        window = np.array(seq[-Win:]).reshape(1, -1)
        y_lit = lit._forward_one_step(O.onehot_pixel_image(window, 6), apply_softmax=False)[0, :, 0, -1]
        if step == 0:
            y_ring = ring.prime(window)[0]
        else:
            y_ring = ring.step(np.array([seq[-1]]))[0]
        np.testing.assert_allclose(y_ring, y_lit, atol=1e-12, err_msg="step %d" % step)
        seq.append(int(np.argmax(y_lit)))


def test_ring_generator_relu_matches_full_window_pass():
    # idea of _tests_/faster_generation/generate.py:27-65: slow path == incremental path
    p = tiny_params()
    rng = np.random.default_rng(9)
    w = O.init_weights(p, rng, np.float64)
    Win = O.input_width(p)
    n = 3
    windows = rng.integers(0, 6, (n, Win))
    ring = O.RingGenerator(p, w, n, head_act="relu", dtype=np.float64)
    got = ring.generate_greedy(windows, 10)
    seqs = [list(r) for r in windows]
    for s in range(10):
        win = np.array([q[-Win:] for q in seqs])
        logits = O.forward_loss(p, w, win, None, dtype=np.float64)["logits"][:, :, -1]
        nxt = np.argmax(logits, axis=1)
        assert np.all(nxt == got[:, s])
        for i in range(n):
            seqs[i].append(int(nxt[i]))


# ---- data codec ---------------------------------------------------------------------------------
def test_mulaw_quirks():
    assert D._MAX["8bit_pcm"] == 128
    # truncating quantiser: 0.0 -> 127, +1.0 -> 255, -1.0 -> 0
    q = D.mulaw_quantize(np.array([0.0, 1.0, -1.0, 0.5, -0.5]))
    assert list(q) == [127, 255, 0, 239, 15]
    # trim always drops the tail sample and all |q-127|<=1 edges
    q = np.array([127, 128, 126, 200, 127, 10, 127, 127], dtype=np.int32)
    assert list(D.trim_silence(q)) == [200, 127]
    # decode: divides by 256, no "-1", q=0 overflows int16
    pcm = D.decode(np.array([128, 192, 255, 64]))
    assert pcm.shape == (4, 2) and np.all(pcm[:, 0] == pcm[:, 1])
    assert pcm[0, 0] == 0                                   # sign(0) == 0
    assert pcm[1, 0] == np.int16(16.0 / 255 * 32768)        # 256**0.5/255, no "-1"
    assert pcm[2, 0] == np.int16((256 ** (255.0 / 256 * 2 - 1)) / 255 * 32768)
    assert pcm[3, 0] == -pcm[1, 0]
    # stereo float path vs mono py2 integer division
    raw = np.array([[1000, 7], [-20000, 7], [32767, 7]], dtype=np.int16)
    st = D.normalize(raw)
    np.testing.assert_allclose(st, raw[:, 0] / 32768.0)
    mono = D.normalize(raw[:, 0])
    assert list(mono) == [0, -1, 0]


@pytest.mark.parametrize("kc,nc", [(1, 2), (3, 2), (2, 3)])
def test_ring_generator_multi_causal_layers_match_full_window(kc, nc):
    # found by the randomised GPU sweep: with causal_conv_filter_width == 1 the history slice must be empty
    p = tiny_params(causal_conv_filter_width=kc, causal_conv_channels=[4, 5, 3][:nc])
    rng = np.random.default_rng(21)
    w = O.init_weights(p, rng, np.float64)
    Win = O.input_width(p)
    windows = rng.integers(0, 6, (2, Win))
    got = O.RingGenerator(p, w, 2, head_act="relu", dtype=np.float64).generate_greedy(windows, 8)
    seqs = [list(r) for r in windows]
    for s in range(8):
        win = np.array([q[-Win:] for q in seqs])
        nxt = np.argmax(O.forward_loss(p, w, win, None, dtype=np.float64)["logits"][:, :, -1], axis=1)
        assert np.all(nxt == got[:, s]), s
        for i in range(2):
            seqs[i].append(int(nxt[i]))
