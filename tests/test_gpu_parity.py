"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle / golden fixtures.

Tolerances are the north star's: FP32 path logits <= 1e-4 max-abs, gradients <= 1e-3
relative, TF32 path logits <= 1e-2 with argmax agreement wherever the oracle's top-2
margin exceeds 2e-2, bit-exact integer work, identical greedy sequences.
"""
import os

import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests.util import make_cfg, make_net, rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

LOGIT_TOL_FP32 = 1e-4
GRAD_TOL = 1e-3
LOGIT_TOL_TF32 = 1e-2


def load_gold(name):
    with np.load(os.path.join(GOLD, name)) as f:
        return {k: f[k] for k in f.files}


def split_gold(gd, prefix):
    return {k[len(prefix):]: v for k, v in gd.items() if k.startswith(prefix)}


def dev(a, dtype=torch.int32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype=dtype)


def run_train_step(net, x, tgt, T):
    """train.py:58-80 through the reference-named methods; returns logits(B,Q,1,T), loss."""
    out = net.forward_causal_block(x)
    out, skip = net.forward_residual_block(out)
    W = out.data.shape[3]
    if W - T >= 1:
        skip = net.slice_1d(skip, W - T)
    logits = net.forward_softmax_block(skip, apply_softmax=False)
    loss = net.cross_entropy(logits, tgt)
    return logits, loss


# ---- golden fixtures: forward, loss, gradients, one optimiser step -------------------------------
@pytest.mark.parametrize("name,T", [("tiny_k2", 20), ("tiny_k3_bias", 61), ("odd", 33)])
def test_train_step_matches_golden(name, T):
    cfg = make_cfg(name)
    gd = load_gold("train_%s.npz" % name)
    net = make_net(cfg, split_gold(gd, "w:"))
    logits, loss = run_train_step(net, gd["x"], gd["target"], T)
    got = logits.data.detach().cpu().numpy()[:, :, 0, :]
    assert got.shape == gd["logits"].shape
    assert np.abs(got - gd["logits"]).max() < LOGIT_TOL_FP32
    assert abs(float(loss.data) - float(gd["loss"])) < 1e-5
    net.update_laerning_rate(1e-3)
    net.backward()
    g = net.get_grads()
    for k, v in split_gold(gd, "g:").items():
        if np.abs(v).max() == 0:
            assert np.abs(g[k]).max() == 0, k
        else:
            assert rel_err(g[k], v) < GRAD_TOL, (k, rel_err(g[k], v))
    net.update()
    assert abs(float(net._norm[0]) - float(gd["norm"])) < 1e-4 * max(1.0, float(gd["norm"]))
    w2 = net.get_weights()
    for k, v in split_gold(gd, "u:").items():
        assert np.abs(w2[k] - v).max() < 2e-5, k


# ---- seeded inputs vs the oracle at the reference's configurations ------------------------------------
@pytest.mark.parametrize("name,B,W,T", [("A", 2, 700, 700), ("B", 1, 600, 343), ("C_small", 2, 1000, 1000),
                                        ("C", 1, 4200, 1129)])
def test_forward_backward_matches_oracle(name, B, W, T):
    cfg = make_cfg(name)
    rng = np.random.default_rng(1234)
    w = O.init_weights(cfg, rng, np.float64)
    x = np.random.default_rng(0).integers(0, cfg.quantization_steps, (B, W)).astype(np.int32)
    tgt = np.random.default_rng(1).integers(0, cfg.quantization_steps, (B, T)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, tgt, train_width=T, dtype=np.float64)
    g_ref = O.backward(cfg, fw)
    net = make_net(cfg, w)
    logits, loss = run_train_step(net, x, tgt, T)
    got = logits.data.detach().cpu().numpy()[:, :, 0, :]
    assert np.abs(got - fw["logits"]).max() < LOGIT_TOL_FP32
    assert abs(float(loss.data) - float(fw["loss"])) < 1e-5
    net.backward()
    g = net.get_grads()
    for k, v in g_ref.items():
        if np.abs(v).max() == 0:
            assert np.abs(g[k]).max() == 0, k
        else:
            assert rel_err(g[k], v) < GRAD_TOL, (k, rel_err(g[k], v))


def test_block_outputs_and_one_hot_input():
    cfg = make_cfg("tiny_k3_bias")
    rng = np.random.default_rng(5)
    w = O.init_weights(cfg, rng, np.float64, bias_scale=0.3)
    x = rng.integers(0, 6, (2, 45)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, None, dtype=np.float64)
    net = make_net(cfg, w)
    onehot = O.onehot_pixel_image(x, 6)                        # the reference's input format
    c = net.forward_causal_block(onehot)
    assert tuple(c.data.shape) == (2, 5, 1, 45)
    assert np.abs(c.data.cpu().numpy()[:, :, 0, :] - fw["causal"]).max() < 1e-5
    out, skip = net.forward_residual_block(c)
    assert np.abs(out.data.cpu().numpy()[:, :, 0, :] - fw["out"]).max() < 1e-5
    assert np.abs(skip.data.cpu().numpy()[:, :, 0, :] - fw["sum_skip"]).max() < 1e-5
    probs = net.forward_one_step(onehot, apply_softmax=True, as_numpy=True)
    want = O.softmax_axis1(fw["logits"])
    assert probs.shape == (2, 6, 1, 45) and np.abs(probs[:, :, 0, :] - want).max() < 1e-5
    # foreign (numpy) inputs to the later blocks are accepted like Chainer accepts raw arrays
    out2, skip2 = net.forward_residual_block(fw["causal"][:, :, None, :].astype(np.float32))
    assert np.abs(skip2.data.cpu().numpy()[:, :, 0, :] - fw["sum_skip"]).max() < 1e-5
    y = net.forward_softmax_block(fw["sum_skip"][:, :, None, -7:].astype(np.float32), apply_softmax=False)
    assert np.abs(y.data.cpu().numpy()[:, :, 0, :] - fw["logits"][:, :, -7:]).max() < 1e-5
    with pytest.raises(Exception, match="width"):
        net.cross_entropy(y, np.zeros((2, 6), np.int32))


def test_full_size_properties_config_c():
    """BASELINE config 2 shape (32 x 16000): size-independent properties instead of the oracle."""
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float32)
    net = make_net(cfg, w)
    B, W = 32, 16000
    x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
    tgt = np.concatenate([x[:, 1:], np.random.default_rng(2).integers(0, 256, (B, 1)).astype(np.int32)], axis=1)
    xd, td = dev(x), dev(tgt)
    net._bind(B, W)
    logits = torch.empty((B, W, 256), dtype=torch.float32, device="cuda")
    from wavenet_b200.wavenet import _ptr, _stream
    from wavenet_b200._lib import check
    check(net._libh.wn_forward_loss(net._h, _ptr(net._params), _ptr(xd), _ptr(td), W, _ptr(net._loss), _ptr(logits),
                                    _stream()))
    loss_full = float(net._loss[0])
    assert abs(loss_full - np.log(256)) < 0.5                  # random init: near-uniform predictions
    # (1) batch items are independent and causal: row 5 alone, truncated to 6000 samples, reproduces its prefix...
    sub = 6000
    fw = O.forward_loss(cfg, w, x[5:6, :sub], None, dtype=np.float32)   # oracle on one short clip (seconds)
    # ...except where the zero prefix (Q1) depends on the width: compare beyond the receptive field of those zeros
    zp_full = [O.zero_prefix(W, 2 ** i, 2) for i in range(10)]
    zp_sub = [O.zero_prefix(sub, 2 ** i, 2) for i in range(10)]
    got = logits[5, :sub].detach().cpu().numpy().T
    if zp_full == zp_sub:
        assert np.abs(got - fw["logits"][0]).max() < LOGIT_TOL_FP32
    else:
        lo = 3 * 1023 + max(max(zp_full), max(zp_sub)) + 2
        assert np.abs(got[:, lo:] - fw["logits"][0][:, lo:]).max() < LOGIT_TOL_FP32
    # (2) loss is the mean of per-row losses: recompute from the returned logits
    lg = logits.double()
    lse = torch.logsumexp(lg, dim=2)
    picked = torch.gather(lg, 2, td.long().unsqueeze(2)).squeeze(2)
    assert abs(float((lse - picked).mean()) - loss_full) < 1e-5
    # (3) gradient of the mean loss w.r.t. the last head bias == mean(softmax - onehot)
    net.backward()
    g = net.get_grads()["softmax_1/b"]
    sm = torch.softmax(lg, dim=2)
    want = (sm.sum(dim=(0, 1)) - torch.bincount(td.flatten().long(), minlength=256).double()) / (B * W)
    assert rel_err(g, want.cpu().numpy()) < GRAD_TOL


def test_adam_multi_step_matches_oracle():
    cfg = make_cfg("odd")
    rng = np.random.default_rng(8)
    w = O.init_weights(cfg, rng, np.float64)
    net = make_net(cfg, w)
    net.update_laerning_rate(1e-3)
    w_ref = {k: v.copy() for k, v in w.items()}
    st = O.new_adam_state(w_ref)
    for step in range(4):
        x = rng.integers(0, 37, (2, 40)).astype(np.int32)
        tgt = rng.integers(0, 37, (2, 25)).astype(np.int32)
        fw = O.forward_loss(cfg, w_ref, x, tgt, train_width=25, dtype=np.float64)
        g_ref = O.backward(cfg, fw)
        O.clip_and_adam(cfg, w_ref, g_ref, st, lr=1e-3)
        loss = net.train_step(dev(x), dev(tgt), train_width=25)
        assert abs(float(loss[0]) - float(fw["loss"])) < 1e-4
    assert net.optimizer.t == 4
    w_got = net.get_weights()
    for k in w_ref:
        assert np.abs(w_got[k] - w_ref[k]).max() < 5e-5, k


# ---- generation ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["tiny_k2", "tiny_k3_bias"])
@pytest.mark.parametrize("act", ["reference", "relu"])
def test_greedy_generation_matches_golden(name, act):
    cfg = make_cfg(name)
    gd = load_gold("gen_%s.npz" % name)
    net = make_net(cfg, split_gold(gd, "w:"), faster=True, head_act=act)
    n, steps = gd["greedy_" + act].shape
    got = net.generate(gd["window"], steps, mode="greedy").cpu().numpy()
    assert np.array_equal(got, gd["greedy_" + act])


@pytest.mark.parametrize("n", [1, 3])
def test_greedy_1000_steps_config_c(n):
    """North star: identical greedy-decoded sequences for 1,000 steps (config C)."""
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
    Win = O.input_width(cfg)
    if n == 1:
        window = np.full((1, Win), 127, dtype=np.int32)                 # generate.py:21
    else:
        window = np.random.default_rng(0).integers(0, 256, (n, Win)).astype(np.int32)
    ring = O.RingGenerator(cfg, w, n, head_act="reference", dtype=np.float64)
    want = ring.generate_greedy(window, 1000)
    net = make_net(cfg, w, faster=True, head_act="reference")
    got = net.generate(window, 1000, mode="greedy").cpu().numpy()
    assert np.array_equal(got, want)


def test_forward_one_step_api_matches_literal_generator():
    """_forward_one_step as generate.py:24-35 drives it, vs the literal rolled-window restatement."""
    cfg = make_cfg("tiny_k2")
    rng = np.random.default_rng(3)
    w = O.init_weights(cfg, rng, np.float64, bias_scale=0.1)
    Win = O.input_width(cfg)
    lit = O.LiteralFastGenerator(cfg, w, np.float64)
    net = make_net(cfg, w, faster=True)
    seq = [127 % 6] * Win
    for step in range(15):
        window = np.array(seq[-Win:], dtype=np.int32).reshape(1, -1)
        onehot = O.onehot_pixel_image(window, 6)
        want = lit._forward_one_step(onehot, apply_softmax=True)[0, :, 0, -1]
        got = net._forward_one_step(onehot, apply_softmax=True, as_numpy=True)
        assert got.shape == (1, 6, 1, 1)
        assert np.abs(got[0, :, 0, -1] - want).max() < 1e-5, step
        seq.append(int(np.argmax(want)))
    net.prev_causal_outputs = None          # cache reset as in _tests_/faster_generation/generate.py:44
    got = net._forward_one_step(O.onehot_pixel_image(np.array(seq[:Win]).reshape(1, -1), 6), as_numpy=True)
    lit2 = O.LiteralFastGenerator(cfg, w, np.float64)
    want = lit2._forward_one_step(O.onehot_pixel_image(np.array(seq[:Win]).reshape(1, -1), 6))[0, :, 0, -1]
    assert np.abs(got[0, :, 0, 0] - want).max() < 1e-5


def test_sampling_follows_softmax_distribution():
    cfg = make_cfg("tiny_k2")
    w = O.init_weights(cfg, np.random.default_rng(4), np.float64)
    Win = O.input_width(cfg)
    n = 4096
    window = np.tile(np.arange(Win, dtype=np.int32) % 6, (n, 1))
    net = make_net(cfg, w, faster=True)
    out = net.generate(window, 1, mode="sample", seed=123).cpu().numpy()[:, 0]
    ring = O.RingGenerator(cfg, w, 1, dtype=np.float64)
    p = O.softmax_axis1(ring.prime(window[:1])[:, :, None])[0, :, 0]
    freq = np.bincount(out, minlength=6) / n
    assert np.abs(freq - p).max() < 4 * np.sqrt(0.25 / n) + 1e-3
    out2 = net.generate(window, 1, mode="sample", seed=123).cpu().numpy()[:, 0]
    assert np.array_equal(out, out2)         # counter-based RNG: same seed, same draw


# ---- data.py on device ---------------------------------------------------------------------------------------
def test_mulaw_device_kernels_bit_exact():
    from oracle import data_oracle as D
    from wavenet_b200 import _lib
    from wavenet_b200.wavenet import _ptr, _stream
    lib = _lib.load()
    rng = np.random.default_rng(0)
    sig = np.concatenate([rng.uniform(-1.2, 1.2, 100000), [0.0, 1.0, -1.0, 0.5, -0.5]])
    want = D.mulaw_quantize(sig)
    sd = torch.from_numpy(sig).cuda()
    q = torch.empty(sig.size, dtype=torch.int32, device="cuda")
    _lib.check(lib.wn_mulaw_encode(_ptr(sd), sig.size, 256, _ptr(q), _stream()))
    got = q.cpu().numpy()
    # bit-exact on every tested input (a mismatch would need the companded value to land within one ulp of a class edge
    # AND libm / CUDA log to differ in that last bit: the reference's own class is decided by libm's last bit there)
    assert np.array_equal(got, want)
    qa = torch.arange(256, dtype=torch.int32, device="cuda")
    out = torch.empty(256, dtype=torch.float64, device="cuda")
    _lib.check(lib.wn_mulaw_decode(_ptr(qa), 256, 256, 32768.0, _ptr(out), _stream()))
    n = (np.arange(256) / 256.0 - 0.5) * 2.0
    ref = np.sign(n) * (256.0 ** np.abs(n)) / 255.0 * 32768.0
    assert np.allclose(out.cpu().numpy(), ref, rtol=1e-12)
    assert np.array_equal(out.cpu().numpy().astype(np.int16)[1:], D.decode(np.arange(256))[1:, 0])


# ---- TF32 tensor-core path (tcgen05) ---------------------------------------------------------------------------
@pytest.mark.parametrize("name,B,W,T", [("C_small", 2, 1000, 1000), ("C", 2, 4200, 1129), ("C", 1, 3071, 1),
                                        ("A", 2, 700, 700), ("B", 1, 600, 343)])
def test_tf32_path_matches_oracle(name, B, W, T):
    """North star: logits within 1e-2 with argmax agreement in the TF32 path.  Argmax must agree wherever
    the oracle's top-2 margin exceeds 2e-2 (closer calls are below the stated logit tolerance)."""
    cfg = make_cfg(name)
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
    x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
    tgt = np.random.default_rng(1).integers(0, 256, (B, T)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, tgt, train_width=T, dtype=np.float64)
    g_ref = O.backward(cfg, fw)
    net = make_net(cfg, w)
    net.set_precision("tf32")
    assert net._libh.wn_tc_active(net._h) == 1, "tcgen05 path must be the one that runs"
    logits, loss = run_train_step(net, x, tgt, T)
    got = logits.data.detach().cpu().numpy()[:, :, 0, :]
    ref = fw["logits"]
    assert np.abs(got - ref).max() < LOGIT_TOL_TF32
    srt = np.sort(ref, axis=1)
    margin = srt[:, -1, :] - srt[:, -2, :]
    agree = got.argmax(axis=1) == ref.argmax(axis=1)
    assert np.all(agree[margin > 2 * LOGIT_TOL_TF32])
    assert agree.mean() > 0.98
    assert abs(float(loss.data) - float(fw["loss"])) < 1e-3
    # gradients: tensor-core forward + exact-fp32 backward kernels on the TF32 tape
    net.backward()
    g = net.get_grads()
    for k, v in g_ref.items():
        if np.abs(v).max() == 0:
            assert np.abs(g[k]).max() == 0, k
        else:
            assert rel_err(g[k], v) < 5e-2, (k, rel_err(g[k], v))   # TF32 forward + TF32 backward


def test_tf32_full_size_agrees_with_fp32_path():
    """BASELINE config 2 shape: TF32 and FP32 GPU paths agree on the loss and logits at 32 x 16000."""
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float32)
    B, W = 32, 16000
    x = np.random.default_rng(0).integers(0, 256, (B, W + 1)).astype(np.int32)
    xd, td = dev(x[:, :W]), dev(x[:, 1:])
    from wavenet_b200.wavenet import _ptr, _stream
    from wavenet_b200._lib import check
    outs = {}
    for prec in ("fp32", "tf32"):
        net = make_net(cfg, w)
        net.set_precision(prec)
        net._bind(B, W)
        logits = torch.empty((B, W, 256), dtype=torch.float32, device="cuda")
        check(net._libh.wn_forward_loss(net._h, _ptr(net._params), _ptr(xd), _ptr(td), W, _ptr(net._loss),
                                        _ptr(logits), _stream()))
        outs[prec] = (float(net._loss[0]), logits)
        del net
    assert abs(outs["fp32"][0] - outs["tf32"][0]) < 1e-3
    err = (outs["fp32"][1] - outs["tf32"][1]).abs().max().item()
    assert err < LOGIT_TOL_TF32, err


@pytest.mark.parametrize("n", [5, 200, 600])
def test_streamed_generator_many_streams(n):
    """Streams-per-CTA variants (1, 2, 4) of the weight-streaming generator vs the oracle, config-C widths."""
    cfg = make_cfg("C_small")
    w = O.init_weights(cfg, np.random.default_rng(7), np.float64)
    Win = O.input_width(cfg)
    window = np.random.default_rng(3).integers(0, 256, (n, Win)).astype(np.int32)
    want = O.RingGenerator(cfg, w, n, head_act="reference", dtype=np.float64).generate_greedy(window, 24)
    net = make_net(cfg, w, faster=True, head_act="reference")
    got = net.generate(window, 24, mode="greedy").cpu().numpy()
    mism = (got != want)
    # fp32 vs fp64 can flip an argmax on a near tie; a flipped stream then diverges: allow a tiny fraction
    bad_streams = mism.any(axis=1).mean()
    assert bad_streams <= 0.01, bad_streams


@pytest.mark.parametrize("n", [1, 5, 13])
def test_cluster_generator_matches_single_cta_kernel_and_continues(n):
    """gen_kernel_v4 (one 8-CTA cluster per stream, up to 15 streams) vs gen_kernel_v3 on the same weights: identical
    greedy sequences; three consecutive wn_gen_run calls continue the state exactly like one long call."""
    import os
    from wavenet_b200 import _lib
    from wavenet_b200._lib import check
    from wavenet_b200.wavenet import _ptr, _stream
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(7), np.float32)

    def run(parts, v4):
        if v4:
            os.environ.pop("WN_GEN_V4", None)
        else:
            os.environ["WN_GEN_V4"] = "0"
        try:
            net = make_net(cfg, w, faster=True)
            window = np.random.default_rng(3).integers(0, 256, (n, O.input_width(cfg))).astype(np.int32)
            net.prime(window)
            outs = []
            for steps in parts:
                out = torch.empty((n, steps), dtype=torch.int32, device="cuda")
                check(net._libh.wn_gen_run(net._gen, _ptr(net._params), steps, _lib.WN_GEN_GREEDY, 0, _ptr(out), _stream()))
                outs.append(out.cpu().numpy())
            return np.concatenate(outs, axis=1)
        finally:
            os.environ.pop("WN_GEN_V4", None)

    a = run([240], True)
    b = run([77, 63, 100], True)
    c = run([240], False)
    assert np.array_equal(a, b)
    assert np.array_equal(a, c)


@pytest.mark.parametrize("n", [16, 20, 37])
def test_many_stream_cluster_generator(n):
    """gen_kernel_v5 (opt-in; 16 streams per 8-CTA cluster; a last, partially filled cluster at n = 20 and 37) at config-C depth:
    greedy sequences against the fp64 ring oracle and against the single-CTA kernel, continuation across wn_gen_run calls,
    and independence of the streams (a shard of the streams generates what it generates inside the full set)."""
    import os
    from wavenet_b200 import _lib
    from wavenet_b200._lib import check
    from wavenet_b200.wavenet import _ptr, _stream
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(7), np.float64)
    window = np.random.default_rng(3).integers(0, 256, (n, O.input_width(cfg))).astype(np.int32)

    def run(win, parts, v5):
        os.environ["WN_GEN_V5"] = "1" if v5 else "0"        # the many-stream cluster kernel is opt-in (slower than v3 today)
        try:
            net = make_net(cfg, w, faster=True, head_act="reference")
            net.prime(win)
            outs = []
            for steps in parts:
                out = torch.empty((win.shape[0], steps), dtype=torch.int32, device="cuda")
                check(net._libh.wn_gen_run(net._gen, _ptr(net._params), steps, _lib.WN_GEN_GREEDY, 0, _ptr(out), _stream()))
                outs.append(out.cpu().numpy())
            return np.concatenate(outs, axis=1)
        finally:
            os.environ.pop("WN_GEN_V5", None)

    steps = 60
    a = run(window, [steps], True)
    want = O.RingGenerator(cfg, w, n, head_act="reference", dtype=np.float64).generate_greedy(window, steps)
    # fp32 vs fp64 may flip an arg-max on a near tie (the stream then diverges): at most one stream
    assert (a != want).any(axis=1).sum() <= 1, (a != want).any(axis=1)
    b = run(window, [17, 30, 13], True)
    assert np.array_equal(a, b)
    c = run(window, [steps], False)                     # single-CTA kernel, same fp32 weights
    assert (a != c).any(axis=1).sum() <= 1
    lo, hi = 3, min(n, 19)                              # a shard that straddles the first cluster boundary when n > 16
    d = run(window[lo:hi], [steps], True)
    assert np.array_equal(d, a[lo:hi])
    e = run(window[lo:hi], [steps], False)              # the default kernels: sharded streams == the same streams unsharded
    assert np.array_equal(e, c[lo:hi])
    os.environ["WN_GEN_NS"] = "2"                       # gen_kernel_v3<2>: the two-streams-per-CTA variant the 256-stream runs use
    try:
        f = run(window, [steps], False)
    finally:
        os.environ.pop("WN_GEN_NS", None)
    assert (f != want).any(axis=1).sum() <= 1 and (f != c).any(axis=1).sum() <= 1


@pytest.mark.parametrize("n,cs", [(24, 8), (130, 8), (130, 4)])
def test_tensor_core_generator(n, cs):
    """gen_kernel_v6 (tcgen05 generator: streams are the MMA M dimension, up to 128 per 8-CTA cluster; automatic from 512
    streams, forced here with WN_GEN_V6=1) at config-C depth: greedy sequences against the fp64 ring oracle and gen_kernel_v3,
    sampled sequences against gen_kernel_v3 (same counter RNG), continuation across wn_gen_run calls (the state moves between
    the fp32 rings and the operand-tile rings), a partially filled second cluster (n = 130), sliced priming, a shard of the
    streams generating what it generates inside the full set, both cluster sizes (8 CTAs: up to 1920 resident streams, 4 CTAs:
    4224), and clusters launched in consecutive waves (what happens beyond the resident capacity)."""
    import os
    from wavenet_b200 import _lib
    from wavenet_b200._lib import check
    from wavenet_b200.faster_wavenet import FasterWaveNet
    from wavenet_b200.wavenet import _ptr, _stream
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(7), np.float64)
    window = np.random.default_rng(3).integers(0, 256, (n, O.input_width(cfg))).astype(np.int32)

    def run(win, parts, v6, mode=_lib.WN_GEN_GREEDY, wave=None):
        os.environ["WN_GEN_V6"] = "1" if v6 else "0"
        os.environ["WN_GEN_V6_CS"] = str(cs)
        if wave:
            os.environ["WN_GEN_V6_WAVE"] = str(wave)
        try:
            net = make_net(cfg, w, faster=True, head_act="reference")
            net.prime(win)
            outs = []
            for i, steps in enumerate(parts):
                out = torch.empty((win.shape[0], steps), dtype=torch.int32, device="cuda")
                check(net._libh.wn_gen_run(net._gen, _ptr(net._params), steps, mode, 11, _ptr(out), _stream()))
                outs.append(out.cpu().numpy())
            return np.concatenate(outs, axis=1)
        finally:
            for k in ("WN_GEN_V6", "WN_GEN_V6_CS", "WN_GEN_V6_WAVE"):
                os.environ.pop(k, None)

    steps = 40
    a = run(window, [steps], True)
    want = O.RingGenerator(cfg, w, n, head_act="reference", dtype=np.float64).generate_greedy(window, steps)
    assert (a != want).any(axis=1).sum() <= 1, (a != want).any(axis=1)      # a near tie may flip one arg-max
    c = run(window, [steps], False)
    assert (a != c).any(axis=1).sum() <= 1
    # 9 + 4 + 27: the middle call is too short for v6 (< 8 steps) and runs on gen_kernel_v3 -> both ring conversions
    b = run(window, [9, 4, 27], True)
    assert (a != b).any(axis=1).sum() <= 1
    b2 = run(window, [13, 27], True)
    assert np.array_equal(a, b2)
    s6 = run(window, [steps], True, _lib.WN_GEN_SAMPLE)
    s3 = run(window, [steps], False, _lib.WN_GEN_SAMPLE)
    assert (s6 != s3).any(axis=1).sum() <= 1
    lo, hi = 3, min(n, 21)
    d = run(window[lo:hi], [steps], True)
    assert np.array_equal(d, a[lo:hi])
    old = FasterWaveNet.PRIME_SLICE
    FasterWaveNet.PRIME_SLICE = 16                                           # priming in slices == priming in one pass
    try:
        e = run(window, [steps], True)
    finally:
        FasterWaveNet.PRIME_SLICE = old
    assert np.array_equal(e, a)
    if n > 128:
        f = run(window, [steps], True, wave=1)                               # two clusters, one per launch
        assert np.array_equal(f, a)


def test_tensor_core_generator_at_resident_capacity():
    """gen_kernel_v6 at the size the bench runs it (every co-resident 4-CTA cluster filled: 4224 streams on a B200, automatic
    kernel choice, fp16x2 priming in slices), through size-independent properties: streams are independent, so (1) streams
    primed with the same window generate the same greedy samples wherever they sit (other cluster, other lane), (2) a handful
    of them agree with the exact-fp32 single-CTA kernel run on just those windows, (3) two launches (waves) equal one."""
    import os
    from wavenet_b200 import _lib
    cfg = make_cfg("C")
    w = O.init_weights(cfg, np.random.default_rng(7), np.float32)
    n = int(_lib.load().wn_gen_mma_capacity(4))
    assert n >= 1024
    base = np.random.default_rng(5).integers(0, 256, (97, O.input_width(cfg))).astype(np.int32)
    window = base[np.arange(n) % 97]
    steps = 24

    def gen(win, env=None):
        import gc
        gc.collect()
        torch.cuda.empty_cache()            # three generators of this size do not fit next to each other (50 GB each)
        os.environ.update(env or {})
        try:
            net = make_net(cfg, w, faster=True, head_act="reference")
            out = net.generate(win, steps, mode="greedy").cpu().numpy()
            del net
            return out
        finally:
            for k in (env or {}):
                os.environ.pop(k, None)
            gc.collect()
            torch.cuda.empty_cache()

    got = gen(window)
    assert np.array_equal(got, got[:97][np.arange(n) % 97])
    ref = gen(base[:12], {"WN_GEN_V6": "0"})
    assert (got[:12] != ref).any(axis=1).sum() <= 1
    waves = gen(window, {"WN_GEN_V6_WAVE": "20"})
    assert np.array_equal(waves, got)


def test_device_crop_batch_matches_reference_create_batch():
    """train_audio/train.py:14-22 restated vs wn_crop_batch with the same np.random stream."""
    rng = np.random.default_rng(0)
    signal = rng.integers(0, 256, 5000).astype(np.int32)
    iw, tw, B = 257, 300, 7
    np.random.seed(3)
    idx = np.random.randint(0, signal.size - tw - iw - 1, size=B)
    want_x = np.stack([signal[s:s + iw + tw] for s in idx])
    want_t = np.stack([signal[s + iw + 1:s + iw + tw + 1] for s in idx])
    cfg = make_cfg("tiny_k2")
    net = make_net(cfg, O.init_weights(cfg, np.random.default_rng(0), np.float64))
    x, t = net.create_batch(dev(signal), idx, iw, tw)
    assert np.array_equal(x.cpu().numpy(), want_x) and np.array_equal(t.cpu().numpy(), want_t)


# ---- seeded randomised sweep: odd shapes, widths shorter than the dilation, T = 1, biases, k = 3 -------------------
def _random_cfg(rng, tc_friendly):
    if tc_friendly:   # multiples of 32 so the tensor-core path is eligible
        R = int(rng.choice([32, 64, 96, 128]))
        G = int(rng.choice([32, 64, 128]))
        S = int(rng.choice([64, 128, 256]))
        Q = int(rng.choice([64, 128, 256]))
        return O.OracleParams(quantization_steps=Q, causal_conv_channels=[R], residual_conv_channels=[G] * int(rng.integers(1, 5)),
                              residual_num_blocks=int(rng.integers(1, 3)), softmax_conv_channels=[S, Q])
    k = int(rng.choice([2, 3]))
    nb = bool(rng.integers(0, 2))
    Q = int(rng.integers(3, 40))
    return O.OracleParams(quantization_steps=Q, causal_conv_channels=[int(rng.integers(1, 20)) for _ in range(int(rng.integers(1, 3)))],
                          causal_conv_filter_width=int(rng.choice([1, 2, 3])), causal_conv_no_bias=nb,
                          residual_conv_filter_width=k, residual_conv_channels=[int(rng.integers(1, 24))] * int(rng.integers(1, 4)),
                          residual_num_blocks=int(rng.integers(1, 3)), residual_conv_dilation_no_bias=bool(rng.integers(0, 2)),
                          residual_conv_projection_no_bias=bool(rng.integers(0, 2)), softmax_conv_no_bias=bool(rng.integers(0, 2)),
                          softmax_conv_channels=[int(rng.integers(1, 30)), int(rng.integers(1, 30)), Q][int(rng.integers(0, 2)):],
                          weight_decay=float(rng.choice([0, 0.01])))


@pytest.mark.parametrize("seed", list(range(8)))
def test_random_shapes_fp32(seed):
    rng = np.random.default_rng(1000 + seed)
    cfg = _random_cfg(rng, False)
    w = O.init_weights(cfg, rng, np.float64, bias_scale=0.2)
    Q = cfg.quantization_steps
    B, W = int(rng.integers(1, 4)), int(rng.integers(1, 70))
    T = int(rng.integers(1, W + 1))
    x = rng.integers(0, Q, (B, W)).astype(np.int32)
    tgt = rng.integers(0, Q, (B, T)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, tgt, train_width=T, dtype=np.float64)
    g_ref = O.backward(cfg, fw)
    net = make_net(cfg, w)
    logits, loss = run_train_step(net, x, tgt, T)
    assert np.abs(logits.data.cpu().numpy()[:, :, 0, :] - fw["logits"]).max() < LOGIT_TOL_FP32
    assert abs(float(loss.data) - float(fw["loss"])) < 1e-5
    net.backward()
    g = net.get_grads()
    for k_, v in g_ref.items():
        if np.abs(v).max() < 1e-12:
            assert np.abs(g[k_]).max() < 1e-6, k_
        else:
            assert rel_err(g[k_], v) < GRAD_TOL, (k_, rel_err(g[k_], v))
    # generator on the same network: greedy steps vs the ring oracle (any shape -> generic kernel)
    n = int(rng.integers(1, 4))
    win = rng.integers(0, Q, (n, O.input_width(cfg))).astype(np.int32)
    if len(cfg.causal_conv_channels) >= 1:
        want = O.RingGenerator(cfg, w, n, head_act="reference", dtype=np.float64).generate_greedy(win, 10)
        gnet = make_net(cfg, w, faster=True)
        got = gnet.generate(win, 10, mode="greedy").cpu().numpy()
        assert (got != want).any(axis=1).mean() <= 0.34    # a near-tie may flip one stream in fp32


@pytest.mark.parametrize("seed", list(range(6)))
def test_random_shapes_tf32(seed):
    rng = np.random.default_rng(2000 + seed)
    cfg = _random_cfg(rng, True)
    w = O.init_weights(cfg, rng, np.float64)
    Q = cfg.quantization_steps
    B, W = int(rng.integers(1, 3)), int(rng.integers(100, 700))
    T = int(rng.integers(1, W + 1))
    x = rng.integers(0, Q, (B, W)).astype(np.int32)
    tgt = rng.integers(0, Q, (B, T)).astype(np.int32)
    fw = O.forward_loss(cfg, w, x, tgt, train_width=T, dtype=np.float64)
    g_ref = O.backward(cfg, fw)
    net = make_net(cfg, w)
    net.set_precision("tf32")
    logits, loss = run_train_step(net, x, tgt, T)
    assert np.abs(logits.data.cpu().numpy()[:, :, 0, :] - fw["logits"]).max() < LOGIT_TOL_TF32
    net.backward()
    g = net.get_grads()
    for k_, v in g_ref.items():
        if np.abs(v).max() < 1e-12:
            assert np.abs(g[k_]).max() < 1e-6, k_
        else:
            assert rel_err(g[k_], v) < 6e-2, (k_, rel_err(g[k_], v), bool(net._libh.wn_tc_active(net._h)))


@pytest.mark.parametrize("name,B,W,T", [("C_small", 2, 1000, 1000), ("C", 3, 2171, 977), ("C", 1, 3200, 3200),
                                        ("C", 9, 16000, 16000), ("C", 5, 12345, 9000)])
def test_tf32_backward_alone_is_tf32_accurate(name, B, W, T):
    """Isolates the tensor-core BACKWARD: same TF32 forward tape, tcgen05 backward vs the exact-fp32 SIMT backward.
    (Against the fp64 oracle the TF32 path shows ~3.5e-2 on gradients, but the exact-fp32 backward run on the same
    TF32 tape shows the same 3.5e-2: at random init the gradient is a small difference of large terms, so the
    7e-4 forward error is amplified ~50x -- conditioning, not backward arithmetic.)  The ragged full-depth case
    (width not a multiple of the 128-row tile, T < W, odd batch) exercises the fused gate-backward + dWp kernel, the
    grouped dzs / dWs launches, the serpentine tile order and the TMA-store clipping of the fused layer kernel; the two
    large cases give every persistent CTA several tiles (1125 and 485 tiles over 148 CTAs), so the barrier phases of all
    rings wrap many times, as they do at the benchmark size."""
    from wavenet_b200 import _lib
    cfg = make_cfg(name)
    w = O.init_weights(cfg, np.random.default_rng(1234), np.float64)
    x = np.random.default_rng(0).integers(0, 256, (B, W)).astype(np.int32)
    tgt = np.random.default_rng(1).integers(0, 256, (B, T)).astype(np.int32)
    net = make_net(cfg, w)
    net.set_precision("tf32")
    run_train_step(net, x, tgt, T)
    assert net._libh.wn_tc_active(net._h) == 1
    net.backward()
    g_tc = net.get_grads()
    _lib.check(net._libh.wn_set_precision(net._h, _lib.WN_PREC_FP32))   # same tape, SIMT backward
    net.backward()
    g_simt = net.get_grads()
    for k, v in g_simt.items():
        if np.abs(v).max() > 0:
            assert rel_err(g_tc[k], v) < 1e-2, (k, rel_err(g_tc[k], v))
        else:
            assert np.abs(g_tc[k]).max() == 0, k


def test_stale_variable_takes_the_external_input_path():
    """A Variable returned by forward_causal_block stands for tape contents; after a second causal pass of the same shape the
    tape holds other data, and the first Variable must go through the external-input path (its .data) instead."""
    cfg = make_cfg("tiny_k2")
    w = O.init_weights(cfg, np.random.default_rng(0), np.float64)
    net = make_net(cfg, w)
    rng = np.random.default_rng(1)
    xa, xb = (rng.integers(0, 6, (2, 40)).astype(np.int32) for _ in range(2))
    ca = net.forward_causal_block(xa)
    net.forward_causal_block(xb)                       # same shape: replaces the tape behind `ca`
    out_a, _ = net.forward_residual_block(ca)
    fresh = make_net(cfg, w)
    want, _ = fresh.forward_residual_block(fresh.forward_causal_block(xa))
    assert np.abs(out_a.data.cpu().numpy() - want.data.cpu().numpy()).max() < 1e-6
