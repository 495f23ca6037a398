"""chainer.functions: the primitives the reference calls (wavenet.py:301,325,330,333,351,360,588,592,613-616;
faster_wavenet.py:108,112).  Arithmetic follows Chainer v2's CPU implementations; dtype follows the inputs."""
import numpy as np

from .function import Function
from .variable import Variable


class _Mul(Function):
    def forward(self, xs):
        return xs[0] * xs[1],

    def backward(self, xs, gy):
        return gy[0] * xs[1], gy[0] * xs[0]


class _Add(Function):
    def forward(self, xs):
        return xs[0] + xs[1],

    def backward(self, xs, gy):
        return gy[0], gy[0]


def _mul(a, b):
    return _Mul()(a, b)


def _add(a, b):
    return _Add()(a, b)


class _Tanh(Function):
    def forward(self, xs):
        self.y = np.tanh(xs[0])
        return self.y,

    def backward(self, xs, gy):
        return gy[0] * (1 - self.y * self.y),


class _Sigmoid(Function):
    def forward(self, xs):
        half = xs[0].dtype.type(0.5)
        self.y = np.tanh(xs[0] * half) * half + half      # chainer/functions/activation/sigmoid.py (v2)
        return self.y,

    def backward(self, xs, gy):
        return gy[0] * self.y * (1 - self.y),


class _ReLU(Function):
    def forward(self, xs):
        return np.maximum(xs[0], 0, dtype=xs[0].dtype),

    def backward(self, xs, gy):
        return gy[0] * (xs[0] > 0),


class _ELU(Function):
    def __init__(self, alpha=1.0):
        self.alpha = alpha

    def forward(self, xs):
        y = xs[0].copy()
        neg = xs[0] < 0
        y[neg] = self.alpha * (np.exp(y[neg]) - 1)
        return y,

    def backward(self, xs, gy):
        gx = gy[0].copy()
        neg = xs[0] < 0
        gx[neg] *= self.alpha * np.exp(xs[0][neg])
        return gx,


class _Reshape(Function):
    def __init__(self, shape):
        self.shape = shape

    def forward(self, xs):
        return xs[0].reshape(self.shape),

    def backward(self, xs, gy):
        return gy[0].reshape(xs[0].shape),


class _Transpose(Function):
    def __init__(self, axes):
        self.axes = axes

    def forward(self, xs):
        return xs[0].transpose(self.axes),

    def backward(self, xs, gy):
        return gy[0].transpose(np.argsort(self.axes)),


class _Softmax(Function):
    def forward(self, xs):
        y = xs[0] - xs[0].max(axis=1, keepdims=True)
        np.exp(y, out=y)
        y /= y.sum(axis=1, keepdims=True)
        self.y = y
        return y,

    def backward(self, xs, gy):
        gx = self.y * gy[0]
        gx -= self.y * gx.sum(axis=1, keepdims=True)
        return gx,


class _SoftmaxCrossEntropy(Function):
    """normalize=True, ignore_label=-1, cache_score=True (Chainer v2 defaults): mean over the valid rows."""

    def forward(self, xs):
        x, t = xs
        m = x.max(axis=1, keepdims=True)
        log_z = m + np.log(np.exp(x - m).sum(axis=1, keepdims=True))
        log_y = x - log_z
        self.y = np.exp(log_y)
        valid = t != -1
        self.count = max(int(valid.sum()), 1)
        picked = log_y[np.arange(t.size), np.maximum(t, 0)] * valid
        return np.asarray(-picked.sum(keepdims=True)[0] / self.count, dtype=x.dtype).reshape(()),

    def backward(self, xs, gy):
        x, t = xs
        gx = self.y.copy()
        gx[np.arange(t.size), np.maximum(t, 0)] -= 1
        gx *= (t != -1)[:, None]
        gx *= gy[0] / self.count
        return gx.astype(x.dtype), None


class _Convolution2D(Function):
    """Cross-correlation, stride 1, no padding (all the reference uses): y[b,o,i,j] = sum W[o,c,p,q] x[b,c,i+p,j+q] + b[o]
    (chainer/functions/connection/convolution_2d.py: im2col + tensordot)."""

    def _cols(self, x, kh, kw):
        return np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(2, 3))   # (B, C, H', W', kh, kw)

    def forward(self, xs):
        x, W = xs[0], xs[1]
        kh, kw = W.shape[2], W.shape[3]
        y = np.tensordot(self._cols(x, kh, kw), W, ((1, 4, 5), (1, 2, 3))).astype(x.dtype, copy=False)   # (B, H', W', O)
        if len(xs) == 3:
            y += xs[2]
        return np.rollaxis(y, 3, 1),

    def backward(self, xs, gys):
        x, W = xs[0], xs[1]
        gy = gys[0]
        kh, kw = W.shape[2], W.shape[3]
        gW = np.tensordot(gy, self._cols(x, kh, kw), ((0, 2, 3), (0, 2, 3))).astype(W.dtype, copy=False)   # (O, C, kh, kw)
        gx = np.zeros_like(x)
        Hp, Wp = gy.shape[2], gy.shape[3]
        for p in range(kh):
            for q in range(kw):
                gx[:, :, p:p + Hp, q:q + Wp] += np.tensordot(gy, W[:, :, p, q], ((1,), (0,))).transpose(0, 3, 1, 2)
        if len(xs) == 3:
            return gx, gW, gy.sum(axis=(0, 2, 3))
        return gx, gW


def tanh(x):
    return _Tanh()(x)


def sigmoid(x):
    return _Sigmoid()(x)


def relu(x):
    return _ReLU()(x)


def elu(x, alpha=1.0):
    return _ELU(alpha)(x)


def reshape(x, shape):
    return _Reshape(shape)(x)


def transpose(x, axes=None):
    return _Transpose(axes)(x)


def softmax(x):
    return _Softmax()(x)


def softmax_cross_entropy(x, t):
    return _SoftmaxCrossEntropy()(x, t)


def convolution_2d(x, W, b=None, stride=1, pad=0):
    assert stride in (1, (1, 1)) and pad in (0, (0, 0)), "the reference only uses stride 1 / pad 0"
    return _Convolution2D()(x, W) if b is None else _Convolution2D()(x, W, b)
