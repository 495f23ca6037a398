"""Minimal NumPy stand-in for the parts of Chainer 2 that /root/reference/{wavenet,faster_wavenet}.py import.

TEST INFRASTRUCTURE ONLY (oracle/): it exists so that the REFERENCE'S OWN SOURCE (mechanically py2->py3-transformed by
oracle/ref_build.py into oracle/_ref/) can be executed in this container, where Chainer / CuPy / Python 2 are absent.
The reference's control flow -- causal pad / reshape / slice (wavenet.py:294-342), the gated unit (:344-368), block
loops (:565-593), cross_entropy (:597-617), backprop through its custom Functions (:202-261), the rolled-window
incremental generator (faster_wavenet.py:50-113) -- then runs for real; only the Chainer PRIMITIVES below are restated
(define-by-run autograd, Convolution2D as cross-correlation, tanh/sigmoid/relu/elu, softmax, softmax_cross_entropy with
normalize=True, reshape/transpose, Adam / WeightDecay / hook order of GradientMethod.update, as published for Chainer v2).
Nothing here is imported by the product (wavenet_b200/).
"""
import contextlib

import numpy as np

from . import cuda  # noqa: F401
from .variable import Variable  # noqa: F401
from . import function  # noqa: F401
from . import functions  # noqa: F401
from . import links  # noqa: F401
from . import optimizer  # noqa: F401
from . import optimizers  # noqa: F401
from . import serializers  # noqa: F401
from .link import Chain, Link  # noqa: F401


class _Config(object):
    train = True
    enable_backprop = True


config = _Config()


@contextlib.contextmanager
def using_config(name, value):
    old = getattr(config, name)
    setattr(config, name, value)
    try:
        yield
    finally:
        setattr(config, name, old)
