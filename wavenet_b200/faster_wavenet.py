"""Host-side mirror of /root/reference/faster_wavenet.py backed by the persistent
generator kernel (wn_gen_* in include/wavenet_b200.h).

Differences from the reference that a caller can observe, all deliberate:
  * `_forward_one_step` returns only the LAST column, shape (n, Q, 1, 1): the
    reference computes the head over the whole window (faster_wavenet.py:105-113)
    but its caller reads `[0, :, 0, -1]` only (train_audio/generate.py:38);
  * n independent streams are accepted (the reference hard-codes batch index 0,
    wavenet.py:286,290,354);
  * `generate()` runs the whole train_audio/generate.py:24-43 loop on the device.

Kernel choice is made by the library (wn_gen.cu): up to 15 streams of the 64/256-channel shape run one 8-CTA
cluster per stream (latency path), more streams one CTA per one or two streams; any other shape uses the generic
kernels.  Arithmetic is exact fp32 in every variant.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check
from .wavenet import WaveNet, Variable, _ptr, _stream


class FasterWaveNet(WaveNet):
    """faster_wavenet.py:11-113."""

    PRIME_SLICE = 256                     # streams per priming pass (the pass keeps ~130 MB of activations per config-C stream)

    def __init__(self, params, seed=None, head_act="reference"):
        WaveNet.__init__(self, params, seed)
        self.head_act = head_act          # "reference": ReLU on the priming call, ELU afterwards (Q2); "relu"
        self.prev_causal_outputs = None   # reset to None to re-prime (_tests_/faster_generation/generate.py:44)
        self._gen = None
        self._gen_n = None
        self._gen_state = None

    def __del__(self):
        try:
            if getattr(self, "_gen", None):
                self._libh.wn_gen_destroy(self._gen)
                self._gen = None
        except Exception:
            pass
        WaveNet.__del__(self)

    @property
    def input_width(self):
        return int(self._libh.wn_input_width(self._h))

    def _ensure_gen(self, n):
        if self._gen is not None and self._gen_n == n:
            return
        if self._gen is not None:
            self._libh.wn_gen_destroy(self._gen)
        g = C.c_void_p()
        check(self._libh.wn_gen_create(self._h, n, 1 if self.head_act == "reference" else 0, C.byref(g)))
        self._gen, self._gen_n = g, n
        nbytes = int(self._libh.wn_gen_state_bytes(g))
        self._gen_state = torch.zeros(nbytes, dtype=torch.uint8, device=self._device)
        check(self._libh.wn_gen_bind_state(g, _ptr(self._gen_state), nbytes))

    def prime(self, window):
        """Priming call (faster_wavenet.py:13-47, 51-52) on (n, Win) samples or a one-hot image."""
        self._need_gpu()
        idx = self._to_indices(window)
        n, W = idx.shape
        if W != self.input_width:
            raise Exception("priming window must be input_width = {} samples wide".format(self.input_width))
        self._ensure_gen(n)
        self._keep["idx"] = idx
        probs = torch.empty((n, self.params.quantization_steps), dtype=torch.float32, device=self._device)
        # the full pass keeps the activation tape of every primed stream: slices of PRIME_SLICE streams bound the workspace
        self._tape_gen = getattr(self, "_tape_gen", 0) + 1      # the priming pass rewrites the tape: older Variables are stale
        for s0 in range(0, n, self.PRIME_SLICE):
            cnt = min(self.PRIME_SLICE, n - s0)
            self._bind(cnt, W)
            check(self._libh.wn_gen_prime_part(self._gen, _ptr(self._params), _ptr(idx[s0:s0 + cnt]), s0, cnt,
                                               _ptr(probs[s0:s0 + cnt]), _stream()))
        self.prev_causal_outputs = True
        return probs

    def step(self, new_samples, apply_softmax=True):
        """One incremental step for int32 samples (n,) -> (n, Q) probabilities / logits."""
        x = new_samples.to(self._device, dtype=torch.int32).contiguous()
        probs = torch.empty((x.shape[0], self.params.quantization_steps), dtype=torch.float32, device=self._device)
        check(self._libh.wn_gen_step(self._gen, _ptr(self._params), _ptr(x), int(bool(apply_softmax)), _ptr(probs),
                                     _stream()))
        return probs

    def _forward_one_step(self, x_batch_data, apply_softmax=True, as_numpy=False):
        """faster_wavenet.py:50-63.  x_batch_data: the current window, one-hot (n,Q,1,Win) or ints (n,Win)."""
        self._need_gpu()
        if not hasattr(self, "prev_causal_outputs") or self.prev_causal_outputs is None:
            probs = self.prime(x_batch_data)
            if not apply_softmax:          # the raw head output of the priming pass (faster_wavenet.py:51-52 -> :17-22)
                check(self._libh.wn_gen_logits(self._gen, _ptr(probs), _stream()))
        else:
            idx = self._to_indices(x_batch_data)
            probs = self.step(idx[:, -1], apply_softmax)
        out = probs.reshape(probs.shape[0], probs.shape[1], 1, 1)
        if as_numpy:
            return out.detach().cpu().numpy()
        return Variable(out)

    def generate(self, window, n_steps, mode="sample", seed=0):
        """generate.py:24-43 on the device: returns int32 (n, n_steps).  mode: 'sample' | 'greedy'."""
        self.prime(window)
        out = torch.empty((self._gen_n, n_steps), dtype=torch.int32, device=self._device)
        m = _lib.WN_GEN_SAMPLE if mode == "sample" else _lib.WN_GEN_GREEDY
        check(self._libh.wn_gen_run(self._gen, _ptr(self._params), int(n_steps), m, int(seed), _ptr(out), _stream()))
        return out
