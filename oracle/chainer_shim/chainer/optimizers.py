"""chainer.optimizers: Adam as in Chainer v2 (the default selected at wavenet.py:83,143); the other names only need to
exist so that the reference's isinstance checks (wavenet.py:483-513) run."""
import math

import numpy as np

from .optimizer import GradientMethod


class Adam(GradientMethod):
    def __init__(self, alpha=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps

    def init_state(self, param, state):
        state["m"] = np.zeros_like(param.data)
        state["v"] = np.zeros_like(param.data)

    @property
    def lr(self):
        fix1 = 1. - self.beta1 ** self.t
        fix2 = 1. - self.beta2 ** self.t
        return self.alpha * math.sqrt(fix2) / fix1

    def update_one_cpu(self, param, state):
        m, v = state["m"], state["v"]
        grad = param.grad
        m += (1 - self.beta1) * (grad - m)
        v += (1 - self.beta2) * (grad * grad - v)
        param.data -= self.lr * m / (np.sqrt(v) + self.eps)


class _Unused(GradientMethod):
    def __init__(self, *a, **k):
        raise NotImplementedError("only Adam is on the hot path")


class AdaGrad(_Unused):
    pass


class AdaDelta(_Unused):
    pass


class NesterovAG(_Unused):
    pass


class RMSprop(_Unused):
    pass


class MomentumSGD(_Unused):
    pass


class SGD(_Unused):
    pass
