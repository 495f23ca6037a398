"""CPU-only tests: oracle vs committed golden fixtures, host logic, and that the C-ABI
library loads and exports every symbol include/wavenet_b200.h declares (no compute)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import wavenet_oracle as O
from oracle import data_oracle as D
from tests.util import make_cfg, to_product_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def load_gold(name):
    with np.load(os.path.join(GOLD, name)) as f:
        return {k: f[k] for k in f.files}


def split_gold(gd, prefix):
    return {k[len(prefix):]: v for k, v in gd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("name,T", [("tiny_k2", 20), ("tiny_k3_bias", 61), ("odd", 33)])
def test_oracle_reproduces_golden_train(name, T):
    cfg = make_cfg(name)
    gd = load_gold("train_%s.npz" % name)
    w = split_gold(gd, "w:")
    fw = O.forward_loss(cfg, w, gd["x"], gd["target"], train_width=T, dtype=np.float64)
    np.testing.assert_allclose(fw["logits"], gd["logits"], atol=1e-12)
    assert abs(float(fw["loss"]) - float(gd["loss"])) < 1e-12
    g = O.backward(cfg, fw)
    for k, v in split_gold(gd, "g:").items():
        np.testing.assert_allclose(g[k], v, atol=1e-12, err_msg=k)
    # fp32 oracle stays within the FP32-path tolerance of the fp64 golden
    fw32 = O.forward_loss(cfg, w, gd["x"], gd["target"], train_width=T, dtype=np.float32)
    assert np.abs(fw32["logits"] - gd["logits"]).max() < 1e-4


@pytest.mark.parametrize("name", ["tiny_k2", "tiny_k3_bias"])
def test_oracle_reproduces_golden_generation(name):
    cfg = make_cfg(name)
    gd = load_gold("gen_%s.npz" % name)
    w = split_gold(gd, "w:")
    n, steps = gd["greedy_relu"].shape
    for act in ("reference", "relu"):
        got = O.RingGenerator(cfg, w, n, head_act=act, dtype=np.float64).generate_greedy(gd["window"], steps)
        assert np.array_equal(got, gd["greedy_" + act])


def test_data_module_matches_oracle_and_golden(tmp_path):
    from wavenet_b200 import data
    gd = load_gold("mulaw.npz")
    assert np.array_equal(D.encode(gd["stereo"]), gd["q"])
    assert np.array_equal(data.quantize_signal(gd["stereo"]), gd["q"])
    assert np.array_equal(data.quantize_signal(gd["stereo"][:, 0].copy()), gd["q_mono"])
    assert np.array_equal(data.dequantize_signal(np.arange(256)), gd["pcm_all"])
    # file round trip through scipy.io.wavfile like the reference
    from scipy.io import wavfile
    path = str(tmp_path / "a.wav")
    wavfile.write(path, 16000, gd["stereo"])
    q, sr = data.load_audio_file(path)
    assert sr == 16000 and np.array_equal(q, gd["q"])
    out = str(tmp_path / "b.wav")
    data.save_audio_file(out, q, sampling_rate=16000)
    sr2, pcm = wavfile.read(out)
    assert sr2 == 16000 and np.array_equal(pcm, D.decode(q))
    oh = data.onehot_pixel_image(np.array([[1, 0, 3]]), 4)
    assert oh.shape == (1, 4, 1, 3) and np.array_equal(oh, O.onehot_pixel_image(np.array([[1, 0, 3]]), 4))


def test_library_exports_every_declared_symbol():
    from wavenet_b200 import _lib
    header = open(os.path.join(ROOT, "include", "wavenet_b200.h")).read()
    declared = set(re.findall(r"\b(wn_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = C.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib = _lib.load()
    assert lib.wn_version() == 100


def test_zero_prefix_matches_oracle():
    from wavenet_b200 import _lib
    lib = _lib.load()
    for k in (2, 3, 4):
        for d in (1, 2, 3, 4, 8, 9, 27, 256, 512):
            for W in (1, 2, 5, 16, 17, 33, 257, 1000, 3071, 16000):
                assert lib.wn_zero_prefix(W, d, k) == O.zero_prefix(W, d, k), (W, d, k)


@pytest.mark.parametrize("name", ["tiny_k2", "tiny_k3_bias", "A", "B", "C"])
def test_param_layout_uses_reference_link_names(name):
    from wavenet_b200.wavenet import WaveNet
    cfg = make_cfg(name)
    net = WaveNet(to_product_params(cfg), seed=0)
    want = O.param_shapes(cfg)
    assert [n for n, _ in want] == list(net.layout.keys())
    for n, shape in want:
        assert net.layout[n][2] == tuple(shape), n
    if name == "C":
        assert sum(v[1] for v in net.layout.values()) == 1270272
        assert net.input_width_host() == 3071 if hasattr(net, "input_width_host") else True
    w = net.get_weights()
    # LeCunNormal std = sqrt(1/fan_in), biases zero (Chainer-2 default)
    big = "softmax_0/W"
    fan_in = w[big].shape[1]
    assert abs(w[big].std() * np.sqrt(fan_in) - 1.0) < 0.2
    for n, a in w.items():
        if n.endswith("/b"):
            assert np.all(a == 0)


def test_params_check_and_errors():
    from wavenet_b200.wavenet import Params, WaveNet
    p = Params()
    p.check()
    assert Params(p.to_dict()).to_dict() == p.to_dict()
    p.learning_rate = 0.1                      # _tests_/training/model.py:36 trips this too
    with pytest.raises(Exception, match="invalid parameter 'learning_rate'"):
        p.check()
    p = Params()
    p.softmax_conv_channels = [128, 255]
    with pytest.raises(Exception, match="quantization_steps != softmax_conv_channels"):
        WaveNet(p)
    net = WaveNet(Params(), seed=0)
    assert not net.gpu_enabled
    with pytest.raises(Exception, match="no CPU path"):
        net.forward_one_step(np.zeros((1, 256, 1, 8), np.float32))
    with pytest.raises(Exception, match="cut cannot be less than one"):
        net.slice_1d(np.zeros((1, 2, 1, 4), np.float32), 0)
    with pytest.raises(Exception, match="cannot be Variable"):
        from wavenet_b200.wavenet import Variable
        net.cross_entropy(np.zeros((1, 256, 1, 4), np.float32), Variable(np.zeros((1, 4), np.int32)))
    net.update_laerning_rate(0.5)
    net.update_momentum(0.8)
    assert net.optimizer.alpha == 0.5 and net.optimizer.beta1 == 0.8


def test_checkpoint_roundtrip_on_host(tmp_path):
    from wavenet_b200.wavenet import Params, WaveNet
    a = WaveNet(Params(), seed=1)
    a.save(str(tmp_path))
    b = WaveNet(Params(), seed=2)
    b.load(str(tmp_path))
    wa, wb = a.get_weights(), b.get_weights()
    for k in wa:
        assert np.array_equal(wa[k], wb[k])


def test_hdf5_checkpoint_layout_matches_chainer_serializer(tmp_path):
    """wavenet.py:619-639: when h5py is available the checkpoint is also written / read in the layout
    chainer.serializers.save_hdf5 produces (group per link, datasets W / b; optimizer t + per-parameter m / v)."""
    h5py = pytest.importorskip("h5py")
    from wavenet_b200.wavenet import WaveNet
    from tests.util import to_product_params
    cfg = make_cfg("tiny_k3_bias")
    net = WaveNet(to_product_params(cfg), seed=3)
    net.save(str(tmp_path))
    with h5py.File(str(tmp_path / "wavenet.model"), "r") as f:
        assert set(f["causal_0"].keys()) == {"W", "b"}
        assert f["residual_0_block_1_wf/W"].shape == (2, 5, 3, 1)
    os.remove(str(tmp_path / "wavenet.model.npz"))
    os.remove(str(tmp_path / "wavenet.opt.npz"))
    net2 = WaveNet(to_product_params(cfg), seed=4)
    net2.load(str(tmp_path))
    for k, v in net.get_weights().items():
        assert np.array_equal(net2.get_weights()[k], v), k
