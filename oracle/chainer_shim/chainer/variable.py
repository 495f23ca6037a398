"""chainer.Variable with define-by-run backward (the subset the reference exercises)."""
import numpy as np


class Variable(object):
    def __init__(self, data, volatile=None, name=None):
        self.data = data
        self.grad = None
        self.creator = None
        self.rank = 0
        self.name = name

    # ---- graph ----
    def backward(self):
        """Reverse-topological walk: every Function's backward runs once, after all of its consumers."""
        if self.grad is None:
            self.grad = np.ones_like(self.data)
        funcs, seen = [], set()

        def add(f):
            if f is not None and id(f) not in seen:
                seen.add(id(f))
                funcs.append(f)

        add(self.creator)
        while funcs:
            funcs.sort(key=lambda f: f.rank)
            f = funcs.pop()
            gys = tuple(o.grad for o in f.outputs)
            gxs = f.backward(tuple(v.data for v in f.inputs), gys)
            if not isinstance(gxs, tuple):
                gxs = (gxs,)
            for v, gx in zip(f.inputs, gxs):
                if gx is None:
                    continue
                v.grad = gx if v.grad is None else v.grad + gx
                add(v.creator)

    def cleargrad(self):
        self.grad = None

    def to_cpu(self):
        pass

    def to_gpu(self):
        pass

    @property
    def shape(self):
        return self.data.shape

    # ---- arithmetic used by the reference: a * b, a + b, 0 + a (sum_skip_connections += z) ----
    def __mul__(self, other):
        from .functions import _mul
        return _mul(self, other)

    def __add__(self, other):
        from .functions import _add
        if not isinstance(other, Variable):
            if np.isscalar(other) and other == 0:
                return self
            other = Variable(np.asarray(other, dtype=self.data.dtype))
        return _add(self, other)

    __radd__ = __add__
