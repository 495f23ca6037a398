"""chainer.optimizer: GradientMethod.update order (Chainer v2): lossfun -> cleargrads -> backward ->
reallocate_cleared_grads -> hooks in insertion order -> t += 1 -> per-parameter update."""
import numpy as np


class WeightDecay(object):
    name = "WeightDecay"

    def __init__(self, rate):
        self.rate = rate

    def __call__(self, opt):
        for p in opt.target.params():
            p.grad += self.rate * p.data


class GradientMethod(object):
    def setup(self, link):
        self.target = link
        self.t = 0
        self._hooks = []
        self._states = {}

    def add_hook(self, hook, name=None):
        self._hooks.append(hook)

    def init_state(self, param, state):
        pass

    def update(self, lossfun=None, *args, **kwds):
        if lossfun is not None:
            loss = lossfun(*args, **kwds)
            self.target.cleargrads()
            loss.backward()
            del loss
        for p in self.target.params():           # reallocate_cleared_grads
            if p.grad is None:
                p.grad = np.zeros_like(p.data)
        for hook in self._hooks:                 # call_hooks
            hook(self)
        self.t += 1
        for p in self.target.params():
            st = self._states.setdefault(id(p), {})
            if not st:
                self.init_state(p, st)
            self.update_one_cpu(p, st)
