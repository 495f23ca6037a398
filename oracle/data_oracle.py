"""CPU oracle for data.py's mu-law codec -- TEST INFRASTRUCTURE ONLY.

Restates /root/reference/data.py:5-58 on arrays (the WAV file I/O itself is
scipy.io.wavfile in both the reference and the product).  PARITY UNPINNED: see
oracle/wavenet_oracle.py header.  Quirks reproduced on purpose:
truncating quantiser, decode divides by Q (not Q-1) and omits the "-1",
`1<<8 - 1` == 128, int16 overflow of q=0 on decode, tail sample always dropped
by the silence trim, and py2 integer division for mono integer WAVs.
"""
import numpy as np

_MAX = {"16bit_pcm": 1 << 15, "32bit_pcm": 1 << 31, "8bit_pcm": 1 << 8 - 1}   # data.py:11-16
_TYPE = {"16bit_pcm": np.int16, "32bit_pcm": np.int32, "8bit_pcm": np.uint8}  # data.py:45-53


def normalize(raw, format="16bit_pcm"):
    """data.py:7-17.  Stereo input keeps channel 0 as float; mono integer input is
    divided in place by an int, which in Python 2 is floor division."""
    signal = np.asarray(raw)
    mx = _MAX[format]
    if signal.ndim > 1:
        signal = signal[:, 0].astype(float)
        return signal / mx
    if np.issubdtype(signal.dtype, np.integer):
        return np.floor_divide(signal.astype(np.int64), mx).astype(signal.dtype)   # py2 `/=` on ints
    return signal / mx


def mulaw_quantize(signal, quantization_steps=256):
    """data.py:19-23 on an already normalised signal."""
    mu = quantization_steps - 1
    signal = np.sign(signal) * np.log(1 + mu * np.absolute(signal)) / np.log(1 + mu)
    return (np.clip(signal * 0.5 + 0.5, 0, 1) * mu).astype(np.int32)


def trim_silence(q):
    """data.py:26-33.  Loop variables keep their last value when no break fires."""
    silence_threshold = 1
    start = 0
    for start in range(q.size):
        if abs(int(q[start]) - 127) > silence_threshold:
            break
    end = 1
    for end in range(1, q.size):
        if abs(int(q[-end]) - 127) > silence_threshold:
            break
    return q[start:-end]


def encode(raw, quantization_steps=256, format="16bit_pcm"):
    """load_audio_file minus the file read (data.py:7-35)."""
    return trim_silence(mulaw_quantize(normalize(raw, format), quantization_steps))


def decode(q, quantization_steps=256, format="16bit_pcm"):
    """save_audio_file minus the file write (data.py:38-57): returns (N, 2) PCM."""
    qf = np.asarray(q).astype(float)
    n = (qf / quantization_steps - 0.5) * 2.0
    mu = quantization_steps - 1
    s = np.sign(n) * ((1 + mu) ** np.absolute(n)) / mu
    s = s * _MAX[format]
    with np.errstate(invalid="ignore", over="ignore"):
        audio = s.reshape((-1, 1)).astype(_TYPE[format])
    return np.repeat(audio, 2, axis=1)
