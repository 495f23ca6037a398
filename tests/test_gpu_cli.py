"""GPU tests of the reference-shaped entry points train_audio/train.py and generate.py (drop-in CLI)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
from scipy.io import wavfile

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small_params(path):
    from wavenet_b200.wavenet import Params
    p = Params()
    p.causal_conv_channels = [64]
    p.residual_conv_channels = [64] * 4
    p.residual_num_blocks = 2
    p.softmax_conv_channels = [256, 256, 256]
    with open(path, "w") as f:
        json.dump(p.to_dict(), f)
    # a checkpoint, so that separate CLI processes see the same weights (model.py:55 loads it at import)
    from wavenet_b200.wavenet import WaveNet
    WaveNet(p, seed=0).save(os.path.dirname(path))


def _run(script, args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, os.path.join(ROOT, "train_audio", script)] + args, cwd=cwd, env=env,
                          capture_output=True, text=True, timeout=600)


def test_generate_cli_fast_and_slow(tmp_path):
    model = tmp_path / "model"
    model.mkdir()
    _small_params(str(model / "wavenet.json"))
    out = tmp_path / "out"
    # --use_faster_wavenet is the README's spelling of --fast (README.md:44 vs train_audio/args.py:14)
    r = _run("generate.py", ["-m", str(model), "-o", str(out), "-s", "0.05", "--use_faster_wavenet", "--seed", "1"], str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    sr, pcm = wavfile.read(str(out / "generated.wav"))
    assert sr == 8000 and pcm.shape == (int(8000 * 0.05) - 1, 2) and pcm.dtype == np.int16
    fast_greedy = _run("generate.py", ["-m", str(model), "-o", str(tmp_path / "g1"), "-s", "0.01", "--fast", "--greedy",
                                       "--precision", "fp32"], str(tmp_path))
    slow_greedy = _run("generate.py", ["-m", str(model), "-o", str(tmp_path / "g2"), "-s", "0.01", "--greedy",
                                       "--precision", "fp32"], str(tmp_path))
    assert fast_greedy.returncode == 0 and slow_greedy.returncode == 0, fast_greedy.stderr[-1500:] + slow_greedy.stderr[-1500:]
    a = wavfile.read(str(tmp_path / "g1" / "generated.wav"))[1]
    b = wavfile.read(str(tmp_path / "g2" / "generated.wav"))[1]
    # the first sample comes from the ReLU priming pass in both paths; later ones differ by design (ELU head, Q2)
    assert a.shape == b.shape and np.array_equal(a[0], b[0])


def test_train_cli_runs_and_saves(tmp_path):
    model = tmp_path / "model"
    model.mkdir()
    _small_params(str(model / "wavenet.json"))
    wav = tmp_path / "wav"
    wav.mkdir()
    n = np.arange(16000)
    sig = (0.5 * np.sin(2 * np.pi * 440 * n / 8000) * 32767).astype(np.int16)
    wavfile.write(str(wav / "tone.wav"), 8000, np.stack([sig, sig], axis=1))
    code = (
        "import sys; sys.argv=['train.py','-w',%r,'-m',%r,'--seed','0'];"
        "import runpy, os; sys.path.insert(0, os.path.join(%r,'train_audio'));"
        "import train; import numpy as np; np.random.seed(0);"
        "train.wavenet.update_laerning_rate(1e-3);"
        "l0 = train.train_audio('tone.wav', batch_size=4, train_width=200, repeat=3);"
        "l1 = train.train_audio('tone.wav', batch_size=4, train_width=200, repeat=3);"
        "print('LOSS', l0, l1); assert np.isfinite(l0) and l1 < l0"
    ) % (str(wav), str(model), ROOT)
    r = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), env=dict(os.environ, PYTHONPATH=ROOT),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1000:] + r.stderr[-2000:]
    assert os.path.isfile(str(model / "wavenet.model.npz")) and os.path.isfile(str(model / "wavenet.opt.npz"))
