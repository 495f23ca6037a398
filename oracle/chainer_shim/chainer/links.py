"""chainer.links.Convolution2D (the only link the reference uses, wavenet.py:263-270,439-440,455)."""
import math

import numpy as np

from . import functions as F
from .link import Link


def _pair(x):
    return x if isinstance(x, (tuple, list)) else (x, x)


class Convolution2D(Link):
    def __init__(self, in_channels, out_channels, ksize, stride=1, pad=0, nobias=False):
        Link.__init__(self)
        kh, kw = _pair(ksize)
        self.stride, self.pad = stride, pad
        # Chainer v2 default initialisers: LeCunNormal for W, zeros for b (parity tests inject their own weights)
        fan_in = in_channels * kh * kw
        self.add_param("W", np.random.normal(0, math.sqrt(1.0 / fan_in), (out_channels, in_channels, kh, kw)).astype(np.float32))
        if nobias:
            self.b = None
        else:
            self.add_param("b", np.zeros(out_channels, dtype=np.float32))

    def __call__(self, x):
        return F.convolution_2d(x, self.W, self.b, self.stride, self.pad)
