"""Pins the oracle to the REFERENCE'S OWN SOURCE.

tests/golden/ref_*.npz were produced by executing /root/reference/{wavenet,faster_wavenet,data}.py themselves
(oracle/ref_build.py: mechanical py2->py3 transform; oracle/chainer_shim: NumPy stand-in for the Chainer primitives) --
see tests/golden/make_ref_golden.py.  The oracle restates the same arithmetic independently (closed-form dilated conv,
manual backward, ring-buffer generator); here both must agree.  The reference computes in float32, the oracle in
float64, hence the 1e-5-class tolerances.  When /root/reference is present (build container) the fixtures are also
regenerated and compared with the committed ones, so that they cannot drift from the sources.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import data_oracle as D
from oracle import wavenet_oracle as O
from tests.util import digest, digest_err, make_cfg, rel_err, weights_from_seed

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

TRAIN_CASES = [("tiny_k2", "tiny_k2"), ("tiny_k3_bias", "tiny_k3_bias"), ("odd", "odd"), ("C_small", "C_small"),
               ("C_small_full", "C_small"), ("B", "B")]


def load(name):
    with np.load(os.path.join(GOLD, name)) as f:
        return {k: f[k] for k in f.files}


@pytest.mark.parametrize("tag,name", TRAIN_CASES)
def test_oracle_matches_reference_source_train_step(tag, name):
    """forward blocks, slice, loss, loss.backward() and one backprop() step (hooks + Adam) of the executed reference
    (wavenet.py:565-617, 515-519, 175-199) vs the oracle."""
    gd = load("ref_train_%s.npz" % tag)
    cfg = make_cfg(name)
    w = {k: v.astype(np.float64) for k, v in weights_from_seed(cfg, int(gd["seed"]), float(gd["bias_scale"])).items()}
    T, tpos, full = int(gd["T"]), gd["tpos"], bool(gd["full"])
    fw = O.forward_loss(cfg, w, gd["x"], gd["target"], train_width=T, dtype=np.float64)
    for key in ("causal", "out", "sum_skip"):
        assert np.abs(fw[key][:, :, tpos] - gd[key]).max() < 2e-5, key
    assert np.abs(fw["logits"] - gd["logits"]).max() < 2e-5
    assert abs(float(fw["loss"]) - float(gd["loss"])) < 2e-6
    g = O.backward(cfg, fw)
    w2 = {k: v.copy() for k, v in w.items()}
    O.clip_and_adam(cfg, w2, {k: v.copy() for k, v in g.items()}, O.new_adam_state(w2), lr=1e-3)
    for k in w:
        if full:
            if np.abs(gd["g:" + k]).max() == 0:
                assert np.abs(g[k]).max() == 0, k           # never reached by backward: Chainer leaves zeros (reallocate_cleared_grads)
            else:
                assert rel_err(g[k], gd["g:" + k]) < 1e-4, (k, rel_err(g[k], gd["g:" + k]))
            assert np.abs(w2[k] - gd["u:" + k]).max() < 3e-6, k
        else:
            if gd["g:" + k][0] == 0:
                assert np.abs(g[k]).max() == 0, k
            else:
                assert digest_err(digest(k, g[k]), gd["g:" + k]) < 1e-4, (k, digest_err(digest(k, g[k]), gd["g:" + k]))
            assert digest_err(digest(k, w2[k] - w[k]), gd["u:" + k]) < 2e-3, k


@pytest.mark.parametrize("tag,name", [("tiny_k2", "tiny_k2"), ("tiny_k3_bias", "tiny_k3_bias"), ("C_small", "C_small")])
def test_oracle_matches_reference_source_generation(tag, name):
    """generate.py:24-43 driven greedily through the executed FasterWaveNet._forward_one_step (faster_wavenet.py:50-113:
    ReLU head on the priming call, ELU head afterwards) and through forward_one_step, vs the oracle's literal
    rolled-window generator, its ring-buffer generator (what the CUDA kernel implements) and its full-window pass."""
    gd = load("ref_gen_%s.npz" % tag)
    cfg = make_cfg(name)
    w = {k: v.astype(np.float64) for k, v in weights_from_seed(cfg, int(gd["seed"]), float(gd["bias_scale"])).items()}
    Q, Win = cfg.quantization_steps, O.input_width(cfg)
    steps = gd["probs_fast"].shape[0]
    ring = O.RingGenerator(cfg, w, 1, head_act="reference", dtype=np.float64)
    assert np.array_equal(ring.generate_greedy(gd["window"], steps)[0], gd["samples_fast"])
    if len(cfg.causal_conv_channels) == 1:       # the literal restatement covers one causal layer
        lit = O.LiteralFastGenerator(cfg, w, np.float64)
        audio = gd["window"][0].copy()
        for s in range(steps):
            p = lit._forward_one_step(O.onehot_pixel_image(audio[-Win:].reshape(1, -1), Q), apply_softmax=True)[0, :, 0, -1]
            assert np.abs(p - gd["probs_fast"][s]).max() < 1e-5, s
            audio = np.append(audio, [np.argmax(p)])
        assert np.array_equal(audio[Win:], gd["samples_fast"])
    audio = gd["window"][0].copy()
    for s in range(steps):                        # slow path: teacher-forced full-window pass each sample
        fw = O.forward_loss(cfg, w, audio[-Win:].reshape(1, -1), None, dtype=np.float64)
        p = O.softmax_axis1(fw["logits"])[0, :, -1]
        assert np.abs(p - gd["probs_slow"][s]).max() < 1e-5, s
        audio = np.append(audio, [np.argmax(p)])
    assert np.array_equal(audio[Win:], gd["samples_slow"])


def test_oracle_matches_reference_source_dilated_conv():
    """DilatedConvolution1D.__call__ (wavenet.py:294-342) and its backward through CausalPadding1d / CausalSlice1d
    (:202-261) as executed, vs the oracle's literal AND closed-form convs (incl. the zero prefix, quirk Q1)."""
    gd = load("ref_dilated_conv.npz")
    i = 0
    while "case%d" % i in gd:
        k, d, W = [int(v) for v in gd["case%d" % i]]
        x, Wt, y = gd["x%d" % i].astype(np.float64), gd["W%d" % i].astype(np.float64), gd["y%d" % i]
        assert np.abs(O.dilated_conv_literal(x, Wt, None, d, k) - y).max() < 1e-5, (k, d, W)
        closed = O.dilated_conv_closed(x[:, :, 0, :], Wt, None, d, k)
        assert np.abs(closed - y[:, :, 0, :]).max() < 1e-5, (k, d, W)
        zp = O.zero_prefix(W, d, k)
        assert np.all(y[:, :, 0, :min(zp, W)] == 0)                      # HARD zeros in the executed reference
        da = gd["gy%d" % i][:, :, 0, :].astype(np.float64).copy()
        da[..., :min(zp, W)] = 0
        dx, gW, _ = O._dconv_backward(x[:, :, 0, :], Wt, False, da, d, k)
        assert np.abs(dx - gd["gx%d" % i][:, :, 0, :]).max() < 1e-4, (k, d, W)
        assert np.abs(gW - gd["gW%d" % i]).max() < 1e-4 * max(1.0, np.abs(gd["gW%d" % i]).max()), (k, d, W)
        i += 1
    assert i == 9


def test_oracle_matches_reference_source_mulaw():
    """data.load_audio_file / save_audio_file / onehot_pixel_image as executed (data.py:5-68), bit for bit."""
    gd = load("ref_mulaw.npz")
    assert np.array_equal(D.encode(gd["stereo"]), gd["q_stereo"])
    assert np.array_equal(D.encode(gd["mono"]), gd["q_mono"])            # py2 integer division: classes {0, 127} only
    assert np.array_equal(D.decode(np.arange(256)), gd["pcm_all"])
    assert np.array_equal(O.onehot_pixel_image(gd["onehot_x"], 256), gd["onehot"])
    from wavenet_b200 import data as PD                                    # the product's host codec, same fixtures
    assert np.array_equal(PD.onehot_pixel_image(gd["onehot_x"], 256), gd["onehot"])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference sources only exist in the build container")
def test_fixtures_regenerate_from_reference_sources(tmp_path):
    """The committed fixtures are what the reference sources produce today (guards against stale fixtures)."""
    env = dict(os.environ, WN_REF_GOLDEN_OUT=str(tmp_path), OMP_NUM_THREADS="4")
    subprocess.run([sys.executable, os.path.join(GOLD, "make_ref_golden.py"), "--only", "tiny_k2,dilated_conv,mulaw"],
                   check=True, env=env, capture_output=True, timeout=600)
    for name in ("ref_train_tiny_k2.npz", "ref_gen_tiny_k2.npz", "ref_dilated_conv.npz", "ref_mulaw.npz"):
        a, b = load(name), None
        with np.load(os.path.join(str(tmp_path), name)) as f:
            b = {k: f[k] for k in f.files}
        assert sorted(a) == sorted(b), name
        for k in a:
            assert np.array_equal(a[k], b[k]), (name, k)
