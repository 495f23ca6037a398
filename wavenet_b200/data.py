"""Host-side mirror of /root/reference/data.py (Python 3): WAV <-> mu-law int32 and
the one-hot image.  The arithmetic is NumPy on the host exactly as in the
reference (it runs once per file, outside the hot path); wn_mulaw_encode/decode
in the C ABI are the on-device variants for resident signals.
"""
import warnings

import numpy as np
from scipy.io import wavfile


def _fmt_max(format):
    if format == "16bit_pcm":
        return 1 << 15
    if format == "32bit_pcm":
        return 1 << 31
    if format == "8bit_pcm":
        return 1 << 8 - 1          # == 128, as in data.py:16 (operator precedence quirk kept)
    raise Exception("unknown format '{}'".format(format))


def quantize_signal(signal, quantization_steps=256, format="16bit_pcm"):
    """data.py:7-33 on an in-memory array (what load_audio_file does after wavfile.read)."""
    # discard R channel to convert to mono if necessary
    if len(signal.shape) > 1:
        signal = signal[:, 0].astype(float)
    max = _fmt_max(format)
    if np.issubdtype(signal.dtype, np.integer):
        # data.py:17 under Python 2: `signal /= max` on an integer array is floor division
        warnings.warn("mono integer PCM: the reference divides the int array in place (Python-2 floor division, data.py:8-17), "
                      "which collapses the signal to classes {0, 127}; kept for parity -- convert the file to stereo or "
                      "float PCM to train on real audio", RuntimeWarning, stacklevel=2)
        signal = np.floor_divide(signal.astype(np.int64), max).astype(signal.dtype)
    else:
        signal = signal / max
    mu = quantization_steps - 1
    signal = np.sign(signal) * np.log(1 + mu * np.absolute(signal)) / np.log(1 + mu)
    quantized_signal = (np.clip(signal * 0.5 + 0.5, 0, 1) * mu).astype(np.int32)
    silence_threshold = 1
    start = 0
    for start in range(quantized_signal.size):
        if abs(int(quantized_signal[start]) - 127) > silence_threshold:
            break
    end = 1
    for end in range(1, quantized_signal.size):
        if abs(int(quantized_signal[-end]) - 127) > silence_threshold:
            break
    return quantized_signal[start:-end]


def load_audio_file(filename, quantization_steps=256, format="16bit_pcm"):
    """data.py:5-35."""
    sampling_rate, signal = wavfile.read(filename)
    return quantize_signal(signal, quantization_steps, format), sampling_rate


def dequantize_signal(quantized_signal, quantization_steps=256, format="16bit_pcm"):
    """data.py:38-57: returns the (N, 2) PCM array save_audio_file writes."""
    quantized_signal = quantized_signal.astype(float)
    normalized_signal = (quantized_signal / quantization_steps - 0.5) * 2.0
    mu = quantization_steps - 1
    signals_1d = np.sign(normalized_signal) * ((1 + mu) ** np.absolute(normalized_signal)) / mu
    max = _fmt_max(format)
    type = {"16bit_pcm": np.int16, "32bit_pcm": np.int32, "8bit_pcm": np.uint8}[format]
    signals_1d = signals_1d * max
    with np.errstate(invalid="ignore", over="ignore"):
        audio = signals_1d.reshape((-1, 1)).astype(type)
    return np.repeat(audio, 2, axis=1)


def save_audio_file(filename, quantized_signal, quantization_steps=256, format="16bit_pcm", sampling_rate=48000):
    """data.py:37-58."""
    wavfile.write(filename, sampling_rate, dequantize_signal(quantized_signal, quantization_steps, format))


def onehot_pixel_image(quantized_signal_batch, quantization_steps=256):
    """data.py:61-68: (B, W) ints -> (B, Q, 1, W) float32.  The B200 backend also accepts
    the (B, W) integers directly and never needs this 256x larger tensor."""
    batchsize = quantized_signal_batch.shape[0]
    width = quantized_signal_batch.shape[1]
    image = np.zeros((batchsize * width, quantization_steps), dtype=np.float32)
    image[np.arange(batchsize * width), quantized_signal_batch.reshape((1, -1))] = 1
    image = image.reshape((batchsize, width, quantization_steps, 1))
    image = image.transpose((0, 2, 3, 1))
    return image
