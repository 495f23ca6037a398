"""CPU oracle for the musyoku/wavenet hot paths -- TEST INFRASTRUCTURE ONLY.

This module is a NumPy restatement of the reference's arithmetic.  It exists to
*check* the CUDA product path (tests/, __graft_entry__.smoke()) and to provide
the host-CPU baseline leg of bench.py.  Nothing under wavenet_b200/ may import
it: the product path has no CPU fallback.

PARITY UNPINNED: the reference (Python 2 + Chainer 2) cannot be imported or run
in this image and its _tests_/ directory holds no golden vectors or assertions
(SURVEY.md section 8c).  The oracle is therefore validated structurally:
  * the literal reshape-trick dilated conv (wavenet.py:294-342) is cross-checked
    against the closed form "causal conv + zero prefix" for many (k, d, W);
  * the manual backward is cross-checked against fp64 finite differences and
    against an independent torch-autograd restatement (tests/test_oracle.py);
  * the all-ones known-input of _tests_/dilated_conv/test_conv.py:13-16 is
    derived by hand;
  * the incremental generator is cross-checked against the full-window pass.

Citations are file:line relative to /root/reference.

Array convention inside the oracle: activations are (B, C, W) -- the
reference's (B, C, 1, W) with the dummy height axis dropped.  Weights keep the
reference's 4-D Chainer shapes and link names (wavenet.py:461-472).
"""
import math
import numpy as np


# --------------------------------------------------------------------------
# hyper-parameters (wavenet.py:100-173)
# --------------------------------------------------------------------------
class OracleParams(object):
    """Attribute bag with the reference's names and defaults (wavenet.py:101-146)."""

    def __init__(self, **kw):
        self.quantization_steps = 256
        self.sampling_rate = 8000
        self.causal_conv_no_bias = True
        self.causal_conv_filter_width = 2
        self.causal_conv_channels = [128]
        self.residual_conv_dilation_no_bias = True
        self.residual_conv_projection_no_bias = True
        self.residual_conv_filter_width = 2
        self.residual_conv_channels = [32] * 9
        self.residual_num_blocks = 2
        self.softmax_conv_no_bias = False
        self.softmax_conv_channels = [128, 256]
        self.optimizer = "adam"
        self.weight_decay = 0
        self.momentum = 0.9
        self.gradient_clipping = 1.0
        for k, v in kw.items():
            if not hasattr(self, k):
                raise Exception("invalid parameter '{}'".format(k))
            setattr(self, k, v)


def config_B():
    """train_audio/model.py:24-43 -- the reference's default network."""
    return OracleParams(causal_conv_channels=[256], residual_conv_channels=[128] * 8,
                        residual_num_blocks=1, softmax_conv_channels=[256, 256])


def config_C():
    """BASELINE.json configs 2-5: 30 layers (d=1..512 x3), 64 residual / 256 skip."""
    return OracleParams(causal_conv_channels=[64], residual_conv_channels=[64] * 10,
                        residual_num_blocks=3, softmax_conv_channels=[256, 256, 256])


def receptive_width(p):
    """train_audio/train.py:36-38."""
    return (p.residual_conv_filter_width ** len(p.residual_conv_channels) - 1) * p.residual_num_blocks + 1


def input_width(p):
    """train_audio/train.py:41-44, generate.py:13-18."""
    return receptive_width(p) + len(p.causal_conv_channels)


# --------------------------------------------------------------------------
# parameter layout (wavenet.py:379-472): reference link names -> 4-D shapes
# --------------------------------------------------------------------------
def param_shapes(p):
    """Ordered list of (name, shape) with Chainer link names (wavenet.py:461-472)."""
    out = []
    kc = p.causal_conv_filter_width
    chans = [(p.quantization_steps, p.causal_conv_channels[0])]
    chans += list(zip(p.causal_conv_channels[:-1], p.causal_conv_channels[1:]))
    for i, (n_in, n_out) in enumerate(chans):            # wavenet.py:391-396
        out.append(("causal_{}/W".format(i), (n_out, n_in, 1, kc)))
        if not p.causal_conv_no_bias:
            out.append(("causal_{}/b".format(i), (n_out,)))
    k = p.residual_conv_filter_width
    R = p.causal_conv_channels[-1]
    S = p.softmax_conv_channels[0]
    for j in range(p.residual_num_blocks):               # wavenet.py:412-444
        for i, G in enumerate(p.residual_conv_channels):
            shape_w = (G, R, 1, k) if i == 0 else (G, R, k, 1)   # wavenet.py:418-424
            base = "residual_{}_block_{}_".format(j, i)
            for nm in ("wf", "wg"):
                out.append((base + nm + "/W", shape_w))
                if not p.residual_conv_dilation_no_bias:
                    out.append((base + nm + "/b", (G,)))
            out.append((base + "projection_block/W", (R, G, 1, 1)))
            if not p.residual_conv_projection_no_bias:
                out.append((base + "projection_block/b", (R,)))
            out.append((base + "projection_softmax/W", (S, G, 1, 1)))
            if not p.residual_conv_projection_no_bias:
                out.append((base + "projection_softmax/b", (S,)))
    hc = list(zip(p.softmax_conv_channels[:-1], p.softmax_conv_channels[1:]))  # wavenet.py:451
    for i, (n_in, n_out) in enumerate(hc):
        out.append(("softmax_{}/W".format(i), (n_out, n_in, 1, 1)))
        if not p.softmax_conv_no_bias:
            out.append(("softmax_{}/b".format(i), (n_out,)))
    return out


def init_weights(p, rng, dtype=np.float32, bias_scale=0.0):
    """Chainer-2 default init: LeCunNormal std=sqrt(1/fan_in), bias 0.

    bias_scale>0 draws non-zero biases so that parity tests exercise bias paths.
    """
    w = {}
    for name, shape in param_shapes(p):
        if name.endswith("/W"):
            fan_in = shape[1] * shape[2] * shape[3]
            w[name] = (rng.standard_normal(shape) * math.sqrt(1.0 / fan_in)).astype(dtype)
        else:
            w[name] = (rng.standard_normal(shape) * bias_scale).astype(dtype)
    return w


# --------------------------------------------------------------------------
# primitive ops
# --------------------------------------------------------------------------
def onehot_pixel_image(q, quantization_steps=256):
    """data.py:61-68: (B, W) ints -> (B, Q, 1, W) float32 one-hot."""
    b, w = q.shape
    img = np.zeros((b * w, quantization_steps), dtype=np.float32)
    img[np.arange(b * w), q.reshape((1, -1))] = 1
    img = img.reshape((b, w, quantization_steps, 1))
    return img.transpose((0, 2, 3, 1))


def causal_padding_1d(x, pad):
    """CausalPadding1d.forward, wavenet.py:216-227 (x is (B, C, H, W))."""
    out = np.zeros(x.shape[:3] + (x.shape[3] + pad,), dtype=x.dtype)
    out[:, :, :, pad:] = x
    return out


def causal_padding_1d_backward(gy, pad):
    """wavenet.py:229-230."""
    return gy[:, :, :, pad:]


def causal_slice_1d(x, cut):
    """CausalSlice1d, wavenet.py:233-254."""
    if cut < 1:
        raise Exception("CausalSlice1d: cut cannot be less than one.")
    return x[:, :, :, cut:]


def causal_slice_1d_backward(x_shape, gy, cut, dtype=np.float32):
    """wavenet.py:256-261."""
    g = np.zeros(x_shape, dtype=dtype)
    g[:, :, :, cut:] = gy
    return g


def conv2d_im2col(x, W, b=None):
    """Chainer-2 CPU Convolution2D forward, stride 1, pad 0: im2col + tensordot.

    x (B, C, H, Wd), W (O, C, kh, kw) -> (B, O, H-kh+1, Wd-kw+1); cross-correlation.
    """
    B, C, H, Wd = x.shape
    O, C2, kh, kw = W.shape
    assert C == C2
    oh, ow = H - kh + 1, Wd - kw + 1
    col = np.empty((B, C, kh, kw, oh, ow), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            col[:, :, i, j] = x[:, :, i:i + oh, j:j + ow]
    y = np.tensordot(col, W, ((1, 2, 3), (1, 2, 3))).astype(x.dtype, copy=False)  # (B, oh, ow, O)
    if b is not None:
        y += b
    return np.rollaxis(y, 3, 1)


def dilated_conv_literal(x4, W, b, dilation, filter_width):
    """DilatedConvolution1D.__call__ restated step by step (wavenet.py:294-342).

    x4 is (B, C, 1, W); returns (B, O, 1, W).
    """
    batchsize, in_ch, _, input_x_width = x4.shape
    out_ch = W.shape[0]
    if dilation == 1:                                         # wavenet.py:298-301
        padded = causal_padding_1d(x4, filter_width - 1)
        return conv2d_im2col(padded, W, b)
    pad = 0
    padded_w = input_x_width
    mod = padded_w % dilation                                 # wavenet.py:308-311
    if mod > 0:
        pad += dilation - mod
        padded_w = input_x_width + pad
    height = padded_w // dilation                             # py2 integer '/', wavenet.py:314
    if height < filter_width:                                 # wavenet.py:315-317
        pad += (filter_width - height) * dilation
        padded_w = input_x_width + pad
    padded = causal_padding_1d(x4, pad) if pad > 0 else x4
    padded = np.ascontiguousarray(padded).reshape((batchsize, in_ch, -1, dilation))  # :325
    out = conv2d_im2col(padded, W, b)                         # :330
    out = np.ascontiguousarray(out).reshape((batchsize, out_ch, 1, -1))              # :333
    cut = out.shape[3] - input_x_width                        # :336
    if cut > 0:
        out = causal_slice_1d(out, cut)
    elif cut < 0:
        out = causal_padding_1d(out, -cut)
    return out


def zero_prefix(width, dilation, filter_width):
    """Number of leading hard-zero outputs of the d>1 branch (quirk Q1).

    Derived from wavenet.py:304-340: pad = (-W mod d) [+ (k-H)d if H<k];
    cut = pad - (k-1)d; cut<0 -> first -cut outputs are zero-padded.
    """
    if dilation == 1:
        return 0
    pad = (-width) % dilation
    height = (width + pad) // dilation
    if height < filter_width:
        pad += (filter_width - height) * dilation
    return max(0, (filter_width - 1) * dilation - pad)


def taps_of(W):
    """Return the k tap matrices (O, C) of a (O,C,1,k) or (O,C,k,1) filter, oldest first."""
    O, C, kh, kw = W.shape
    if kh == 1:
        return [W[:, :, 0, i] for i in range(kw)]
    return [W[:, :, i, 0] for i in range(kh)]


def shift_right(x, s):
    """y[..., t] = x[..., t-s] with zero fill (x is (B, C, W))."""
    if s == 0:
        return x
    y = np.zeros_like(x)
    if s < x.shape[-1]:
        y[..., s:] = x[..., :-s]
    return y


def shift_left(x, s):
    """y[..., t] = x[..., t+s] with zero fill."""
    if s == 0:
        return x
    y = np.zeros_like(x)
    if s < x.shape[-1]:
        y[..., :-s] = x[..., s:]
    return y


def dilated_conv_closed(x, W, b, dilation, filter_width):
    """Closed form of wavenet.py:294-342 on (B, C, W) arrays.

    a[:, t] = sum_i W_i x[:, t-(k-1-i)d] (+b) for t >= zp, hard 0 for t < zp.
    """
    k = filter_width
    taps = taps_of(W)
    Wd = x.shape[-1]
    acc = None
    for i in range(k):
        xs = shift_right(x, (k - 1 - i) * dilation)
        term = np.einsum("oc,bcw->bow", taps[i], xs, optimize=True)
        acc = term if acc is None else acc + term
    if b is not None:
        acc = acc + b[None, :, None]
    zp = zero_prefix(Wd, dilation, k)
    if zp > 0:
        acc[..., :min(zp, Wd)] = 0
    return acc.astype(x.dtype, copy=False)


def sigmoid(x):
    """Chainer F.sigmoid CPU forward: tanh(x/2)/2 + 1/2."""
    half = x.dtype.type(0.5)
    return np.tanh(x * half) * half + half


def elu(x):
    """F.elu alpha=1 (faster_wavenet.py:108)."""
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0))).astype(x.dtype, copy=False)


def conv1x1(x, W, b):
    y = np.einsum("oc,bcw->bow", W[:, :, 0, 0], x, optimize=True)
    if b is not None:
        y = y + b[None, :, None]
    return y.astype(x.dtype, copy=False)


def softmax_axis1(y):
    m = y.max(axis=1, keepdims=True)
    e = np.exp(y - m)
    return e / e.sum(axis=1, keepdims=True)


# --------------------------------------------------------------------------
# network forward (wavenet.py:556-617), closed form, with caches for backward
# --------------------------------------------------------------------------
def _layers(p):
    k = p.residual_conv_filter_width
    for j in range(p.residual_num_blocks):
        for i in range(len(p.residual_conv_channels)):
            yield j, i, k ** i, "residual_{}_block_{}_".format(j, i)


def forward_causal_block(p, w, x_onehot):
    """wavenet.py:565-570 on (B, Q, W) one-hot; no activation between layers (Q4)."""
    h = x_onehot
    cache = []
    kc = p.causal_conv_filter_width
    for i in range(len(p.causal_conv_channels)):
        cache.append(h)
        h = dilated_conv_closed(h, w["causal_{}/W".format(i)], w.get("causal_{}/b".format(i)), 1, kc)
    return h, cache


def forward_residual_block(p, w, x):
    """wavenet.py:572-582 + ResidualConvLayer.__call__ :358-368."""
    k = p.residual_conv_filter_width
    sum_skip = 0
    cache = []
    for j, i, d, base in _layers(p):
        a_f = dilated_conv_closed(x, w[base + "wf/W"], w.get(base + "wf/b"), d, k)
        a_g = dilated_conv_closed(x, w[base + "wg/W"], w.get(base + "wg/b"), d, k)
        tf = np.tanh(a_f)
        sg = sigmoid(a_g)
        z = tf * sg
        out = conv1x1(z, w[base + "projection_block/W"], w.get(base + "projection_block/b")) + x
        skip = conv1x1(z, w[base + "projection_softmax/W"], w.get(base + "projection_softmax/b"))
        sum_skip = sum_skip + skip
        cache.append((x, tf, sg))
        x = out
    return x, sum_skip, cache


def forward_softmax_block(p, w, y, apply_softmax=True, act="relu"):
    """wavenet.py:584-593 (act='relu') / faster_wavenet.py:105-113 (act='elu')."""
    cache = []
    for i in range(len(p.softmax_conv_channels) - 1):
        cache.append(y)
        u = np.maximum(y, 0) if act == "relu" else elu(y)
        y = conv1x1(u, w["softmax_{}/W".format(i)], w.get("softmax_{}/b".format(i)))
    if apply_softmax:
        y = softmax_axis1(y)
    return y, cache


def cross_entropy(logits, target):
    """wavenet.py:597-617: F.softmax_cross_entropy, mean over B*T rows.

    logits (B, Q, T), target int (B, T).  Returns (loss, dlogits).
    """
    if logits.shape[2] != target.shape[1]:
        raise Exception("raw_network_output.width != target.width")
    B, Q, T = logits.shape
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m)
    s = e.sum(axis=1, keepdims=True)
    logp = logits - m - np.log(s)
    bi = np.arange(B)[:, None]
    ti = np.arange(T)[None, :]
    picked = logp[bi, target, ti]
    n = B * T
    loss = -(picked.sum(dtype=np.float64)) / n
    d = e / s
    d[bi, target, ti] -= 1
    d = d / logits.dtype.type(n)
    return logits.dtype.type(loss), d.astype(logits.dtype, copy=False)


def forward_loss(p, w, x_idx, target, train_width=None, dtype=np.float32):
    """Full teacher-forced forward + loss.

    train_width=None: forward_one_step(apply_softmax=False) + cross_entropy over
    the full width (wavenet.py:556,597).  train_width=T: train.py:66-78 -- the
    head and loss see only the last T columns (slice_1d, train.py:72-73).
    Returns dict with logits (B,Q,T), loss, caches.
    """
    B, W = x_idx.shape
    Q = p.quantization_steps
    x1h = onehot_pixel_image(x_idx, Q)[:, :, 0, :].astype(dtype)
    w = {k_: v.astype(dtype, copy=False) for k_, v in w.items()}
    h, c_cache = forward_causal_block(p, w, x1h)
    out, sum_skip, r_cache = forward_residual_block(p, w, h)
    T = W if train_width is None else train_width
    y = sum_skip[:, :, W - T:]
    logits, h_cache = forward_softmax_block(p, w, y, apply_softmax=False)
    res = dict(causal=h, out=out, sum_skip=sum_skip, logits=logits, T=T,
               c_cache=c_cache, r_cache=r_cache, h_cache=h_cache, w=w, x_idx=x_idx)
    if target is not None:
        loss, dlogits = cross_entropy(logits, target)
        res["loss"] = loss
        res["dlogits"] = dlogits
    return res


# --------------------------------------------------------------------------
# backward (what Chainer autograd computes for train.py:80 / wavenet.py:515-519)
# --------------------------------------------------------------------------
def _set_tap_grad(gW, i, val):
    if gW.shape[2] == 1:
        gW[:, :, 0, i] = val
    else:
        gW[:, :, i, 0] = val


def _dconv_backward(x, W, has_b, da, d, k):
    """Backward of dilated_conv_closed.  da must already be zero for t < zp."""
    gW = np.zeros_like(W)
    taps = taps_of(W)
    dx = np.zeros_like(x)
    for i in range(k):
        s = (k - 1 - i) * d
        xs = shift_right(x, s)
        _set_tap_grad(gW, i, np.einsum("bow,bcw->oc", da, xs, optimize=True))
        dx += shift_left(np.einsum("oc,bow->bcw", taps[i], da, optimize=True), s)
    gb = da.sum(axis=(0, 2)) if has_b else None
    return dx, gW, gb


def backward(p, fw):
    """Gradients of fw['loss'] w.r.t. every parameter; dict name -> array.

    Parameters not reached by backward (the last layer's projection_block) get
    zero gradients, as Chainer's reallocate_cleared_grads does.
    """
    w = fw["w"]
    k = p.residual_conv_filter_width
    g = {}
    dy = fw["dlogits"]
    n_head = len(p.softmax_conv_channels) - 1
    for i in reversed(range(n_head)):
        y_prev = fw["h_cache"][i]
        u = np.maximum(y_prev, 0)
        Wn = "softmax_{}/W".format(i)
        g[Wn] = np.einsum("bow,bcw->oc", dy, u, optimize=True)[:, :, None, None]
        if ("softmax_{}/b".format(i)) in w:
            g["softmax_{}/b".format(i)] = dy.sum(axis=(0, 2))
        du = np.einsum("oc,bow->bcw", w[Wn][:, :, 0, 0], dy, optimize=True)
        dy = du * (y_prev > 0)
    B, S, T = dy.shape
    Wd = fw["sum_skip"].shape[2]
    dskip = np.zeros((B, S, Wd), dtype=dy.dtype)
    dskip[:, :, Wd - T:] = dy                                  # backward of slice_1d (wavenet.py:256-261)
    dout = np.zeros_like(fw["causal"])                         # final 'output' is unused (train.py:72)
    layers = list(_layers(p))
    for li in reversed(range(len(layers))):
        j, i, d, base = layers[li]
        x, tf, sg = fw["r_cache"][li]
        z = tf * sg
        Wp = w[base + "projection_block/W"]
        Ws = w[base + "projection_softmax/W"]
        g[base + "projection_block/W"] = np.einsum("bow,bcw->oc", dout, z, optimize=True)[:, :, None, None]
        g[base + "projection_softmax/W"] = np.einsum("bow,bcw->oc", dskip, z, optimize=True)[:, :, None, None]
        if (base + "projection_block/b") in w:
            g[base + "projection_block/b"] = dout.sum(axis=(0, 2))
            g[base + "projection_softmax/b"] = dskip.sum(axis=(0, 2))
        dz = np.einsum("oc,bow->bcw", Wp[:, :, 0, 0], dout, optimize=True) \
            + np.einsum("oc,bow->bcw", Ws[:, :, 0, 0], dskip, optimize=True)
        one = x.dtype.type(1)
        da_f = dz * sg * (one - tf * tf)
        da_g = dz * tf * sg * (one - sg)
        zp = zero_prefix(Wd, d, k)
        if zp > 0:
            da_f[..., :min(zp, Wd)] = 0
            da_g[..., :min(zp, Wd)] = 0
        has_b = (base + "wf/b") in w
        dx_f, g[base + "wf/W"], gbf = _dconv_backward(x, w[base + "wf/W"], has_b, da_f, d, k)
        dx_g, g[base + "wg/W"], gbg = _dconv_backward(x, w[base + "wg/W"], has_b, da_g, d, k)
        if has_b:
            g[base + "wf/b"] = gbf
            g[base + "wg/b"] = gbg
        dout = dout + dx_f + dx_g
    kc = p.causal_conv_filter_width
    dh = dout
    for i in reversed(range(len(p.causal_conv_channels))):
        name = "causal_{}/W".format(i)
        has_b = ("causal_{}/b".format(i)) in w
        dh, g[name], gb = _dconv_backward(fw["c_cache"][i], w[name], has_b, dh, 1, kc)
        if has_b:
            g["causal_{}/b".format(i)] = gb
    for name in w:
        g[name] = g[name].astype(w[name].dtype, copy=False)
    return g


# --------------------------------------------------------------------------
# optimiser glue (wavenet.py:175-199, 457-519; Chainer-2 Adam)
# --------------------------------------------------------------------------
def clip_and_adam(p, w, g, state, lr, beta2=0.999, eps=1e-8):
    """One optimizer.update(): [WeightDecay] -> GradientClipping -> Adam, in place.

    state = {'t': int, 'm': {name: arr}, 'v': {name: arr}}.
    Returns the global gradient norm before clipping.
    """
    names = [n for n, _ in param_shapes(p)]
    if p.weight_decay > 0:                                    # wavenet.py:477-478
        for n in names:
            g[n] = g[n] + w[n].dtype.type(p.weight_decay) * w[n]
    sq = 0.0
    for n in names:                                           # sum_sqnorm, wavenet.py:175-182
        x = g[n].ravel()
        sq += float(x.dot(x))
    norm = math.sqrt(sq)
    if p.gradient_clipping > 0 and norm != 0:                 # wavenet.py:190-199
        rate = p.gradient_clipping / norm
        if rate < 1:
            for n in names:
                g[n] = g[n] * g[n].dtype.type(rate)
    state["t"] += 1
    t = state["t"]
    b1 = p.momentum
    fix1 = 1.0 - b1 ** t
    fix2 = 1.0 - beta2 ** t
    step = lr * math.sqrt(fix2) / fix1
    for n in names:
        dt = w[n].dtype.type
        m, v = state["m"][n], state["v"][n]
        m += dt(1 - b1) * (g[n] - m)
        v += dt(1 - beta2) * (g[n] * g[n] - v)
        w[n] -= dt(step) * m / (np.sqrt(v) + dt(eps))
    return norm


def new_adam_state(w):
    return dict(t=0, m={n: np.zeros_like(a) for n, a in w.items()},
                v={n: np.zeros_like(a) for n, a in w.items()})


# --------------------------------------------------------------------------
# literal (reference-shaped) forward used as CPU baseline and cross-check
# --------------------------------------------------------------------------
def forward_literal(p, w, x_onehot4, apply_softmax=False, act="relu"):
    """wavenet.py:556-593 with the reference's op sequence on (B,C,1,W) arrays:
    one-hot input, pad copies, reshape trick, im2col+tensordot convs."""
    kc = p.causal_conv_filter_width
    k = p.residual_conv_filter_width
    h = x_onehot4
    for i in range(len(p.causal_conv_channels)):
        h = dilated_conv_literal(h, w["causal_{}/W".format(i)], w.get("causal_{}/b".format(i)), 1, kc)
    sum_skip = 0
    x = h
    outs = []
    for j, i, d, base in _layers(p):
        a_f = dilated_conv_literal(x, w[base + "wf/W"], w.get(base + "wf/b"), d, k)
        a_g = dilated_conv_literal(x, w[base + "wg/W"], w.get(base + "wg/b"), d, k)
        z = np.tanh(a_f) * sigmoid(a_g)
        out = conv2d_im2col(z, w[base + "projection_block/W"], w.get(base + "projection_block/b")) + x
        skip = conv2d_im2col(z, w[base + "projection_softmax/W"], w.get(base + "projection_softmax/b"))
        sum_skip = sum_skip + skip
        outs.append((out, skip))
        x = out
    y = sum_skip
    for i in range(len(p.softmax_conv_channels) - 1):
        u = np.maximum(y, 0) if act == "relu" else elu(y)
        y = conv2d_im2col(u, w["softmax_{}/W".format(i)], w.get("softmax_{}/b".format(i)))
    if apply_softmax:
        y = softmax_axis1(y)
    return dict(causal=h, out=x, sum_skip=sum_skip, y=y, layer_outs=outs)


# --------------------------------------------------------------------------
# incremental generator (faster_wavenet.py)
# --------------------------------------------------------------------------
class LiteralFastGenerator(object):
    """FasterWaveNet restated literally: rolled full-window caches, head over the
    whole window, ReLU on the priming call and ELU afterwards (quirk Q2)."""

    def __init__(self, p, w, dtype=np.float32):
        self.p = p
        self.w = {k_: v.astype(dtype) for k_, v in w.items()}
        self.dtype = dtype
        self.prev_causal_outputs = None
        self.prev_residual_outputs = None

    def _conv_step(self, W, b, window, d, k):
        """DilatedConvolution1D._forward, wavenet.py:281-292 (window (1,C,1,Win))."""
        taps = taps_of(W)
        acc = 0
        for n in range(k):
            acc = acc + taps[k - 1 - n].dot(window[0, :, 0, -d * n - 1])
        if b is not None:
            acc = acc + b
        return acc.astype(self.dtype)

    def forward_one_step(self, x4, apply_softmax=True):
        """faster_wavenet.py:13-47 (priming: full pass, caches every layer)."""
        r = forward_literal(self.p, self.w, x4.astype(self.dtype), apply_softmax, act="relu")
        # faster_wavenet.py:29 -- only the output of each causal layer is cached; with one
        # causal layer that is r['causal'].
        assert len(self.p.causal_conv_channels) == 1, "literal generator: single causal layer only"
        self.prev_causal_outputs = [r["causal"].copy()]
        self.prev_residual_outputs = [[o.copy(), s.copy()] for (o, s) in r["layer_outs"]]
        return r["y"]

    def _forward_one_step(self, x4, apply_softmax=True):
        """faster_wavenet.py:50-63."""
        if self.prev_causal_outputs is None:
            return self.forward_one_step(x4, apply_softmax)
        p, w = self.p, self.w
        kc, k = p.causal_conv_filter_width, p.residual_conv_filter_width
        inp = x4.astype(self.dtype)
        for i in range(len(p.causal_conv_channels)):              # :65-78
            o = self._conv_step(w["causal_{}/W".format(i)], w.get("causal_{}/b".format(i)), inp, 1, kc)
            prev = np.roll(self.prev_causal_outputs[i], -1, axis=3)
            prev[0, :, 0, -1] = o
            self.prev_causal_outputs[i] = prev
            inp = prev
        sum_skip = 0
        for li, (j, i, d, base) in enumerate(_layers(p)):           # :80-103
            a_f = self._conv_step(w[base + "wf/W"], w.get(base + "wf/b"), inp, d, k)
            a_g = self._conv_step(w[base + "wg/W"], w.get(base + "wg/b"), inp, d, k)
            z = np.tanh(a_f) * sigmoid(a_g)                         # wavenet.py:351
            pb = w[base + "projection_block/W"][:, :, 0, 0].dot(z)
            ps = w[base + "projection_softmax/W"][:, :, 0, 0].dot(z)
            if (base + "projection_block/b") in w:
                pb = pb + w[base + "projection_block/b"]
                ps = ps + w[base + "projection_softmax/b"]
            o = pb + inp[0, :, 0, -1]                               # wavenet.py:354
            prev_o, prev_z = self.prev_residual_outputs[li]
            prev_o = np.roll(prev_o, -1, axis=3)
            prev_o[0, :, 0, -1] = o
            prev_z = np.roll(prev_z, -1, axis=3)
            prev_z[0, :, 0, -1] = ps
            self.prev_residual_outputs[li] = [prev_o, prev_z]
            sum_skip = sum_skip + prev_z
            inp = prev_o
        y = sum_skip                                                 # :105-113, ELU head
        for i in range(len(p.softmax_conv_channels) - 1):
            y = conv2d_im2col(elu(y), w["softmax_{}/W".format(i)], w.get("softmax_{}/b".format(i)))
        if apply_softmax:
            y = softmax_axis1(y)
        return y


class RingGenerator(object):
    """Same arithmetic as LiteralFastGenerator for the last column, restated with
    per-layer dilation ring buffers (what the CUDA generator implements), batched
    over independent streams.  head_act: 'reference' (ReLU on the priming step,
    ELU afterwards, Q2) | 'relu' | 'elu'."""

    def __init__(self, p, w, n_streams, head_act="reference", dtype=np.float32):
        self.p, self.n, self.dtype, self.head_act = p, n_streams, dtype, head_act
        self.w = {k_: v.astype(dtype) for k_, v in w.items()}
        self.primed = False

    def prime(self, window_idx):
        """Full pass over (n, Win) int windows (faster_wavenet.py:13-47); returns last-column logits."""
        p = self.p
        fw = forward_loss(p, self.w, window_idx, None, dtype=self.dtype)
        k, kc = p.residual_conv_filter_width, p.causal_conv_filter_width
        nc = len(p.causal_conv_channels)
        # history needed by the next step: the last (k-1)*d columns of every conv input
        self.idx_hist = window_idx[:, window_idx.shape[1] - (kc - 1):].copy()
        Wn = window_idx.shape[1]
        self.causal_hist = [fw["c_cache"][i][:, :, Wn - (kc - 1):].copy() for i in range(1, nc)]   # empty when kc == 1
        self.rings = []
        for li, (j, i, d, base) in enumerate(_layers(p)):
            x = fw["r_cache"][li][0]
            need = (k - 1) * d
            ring = np.zeros((self.n, x.shape[1], need), dtype=self.dtype)
            have = min(need, x.shape[2])
            ring[:, :, need - have:] = x[:, :, x.shape[2] - have:]
            self.rings.append(ring)
        self.primed = True
        act = "relu" if self.head_act in ("reference", "relu") else "elu"
        y, _ = forward_softmax_block(p, self.w, fw["sum_skip"][:, :, -1:], apply_softmax=False, act=act)
        return y[:, :, 0]

    def step(self, new_idx):
        """One incremental step for new samples (n,) ; returns logits (n, Q)."""
        p, w = self.p, self.w
        k, kc = p.residual_conv_filter_width, p.causal_conv_filter_width
        nc = len(p.causal_conv_channels)
        # causal layer 0 on one-hot == column gather (wavenet.py:281-286)
        idx_win = np.concatenate([self.idx_hist, new_idx[:, None]], axis=1)      # (n, kc)
        W0 = w["causal_0/W"]
        h = 0
        for jn in range(kc):
            h = h + W0[:, idx_win[:, jn], 0, jn].T
        if "causal_0/b" in w:
            h = h + w["causal_0/b"]
        h = h.astype(self.dtype)
        self.idx_hist = idx_win[:, 1:]
        for i in range(1, nc):
            hist = np.concatenate([self.causal_hist[i - 1], h[:, :, None]], axis=2)   # (n, C, kc)
            Wi = w["causal_{}/W".format(i)]
            h2 = 0
            for jn in range(kc):
                h2 = h2 + hist[:, :, jn].dot(Wi[:, :, 0, jn].T)
            if ("causal_{}/b".format(i)) in w:
                h2 = h2 + w["causal_{}/b".format(i)]
            self.causal_hist[i - 1] = hist[:, :, 1:]
            h = h2.astype(self.dtype)
        x = h
        sum_skip = 0
        for li, (j, i, d, base) in enumerate(_layers(p)):
            ring = self.rings[li]
            taps_f, taps_g = taps_of(w[base + "wf/W"]), taps_of(w[base + "wg/W"])
            a_f = x.dot(taps_f[k - 1].T)
            a_g = x.dot(taps_g[k - 1].T)
            for n_ in range(1, k):
                past = ring[:, :, ring.shape[2] - n_ * d]
                a_f = a_f + past.dot(taps_f[k - 1 - n_].T)
                a_g = a_g + past.dot(taps_g[k - 1 - n_].T)
            if (base + "wf/b") in w:
                a_f = a_f + w[base + "wf/b"]
                a_g = a_g + w[base + "wg/b"]
            z = np.tanh(a_f) * sigmoid(a_g)
            pb = z.dot(w[base + "projection_block/W"][:, :, 0, 0].T)
            ps = z.dot(w[base + "projection_softmax/W"][:, :, 0, 0].T)
            if (base + "projection_block/b") in w:
                pb = pb + w[base + "projection_block/b"]
                ps = ps + w[base + "projection_softmax/b"]
            out = (pb + x).astype(self.dtype)
            sum_skip = sum_skip + ps
            self.rings[li] = np.concatenate([ring[:, :, 1:], x[:, :, None]], axis=2)
            x = out
        y = sum_skip.astype(self.dtype)
        act = "relu" if self.head_act == "relu" else "elu"
        for i in range(len(p.softmax_conv_channels) - 1):
            u = np.maximum(y, 0) if act == "relu" else elu(y)
            y = u.dot(w["softmax_{}/W".format(i)][:, :, 0, 0].T)
            if ("softmax_{}/b".format(i)) in w:
                y = y + w["softmax_{}/b".format(i)]
        return y.astype(self.dtype)

    def generate_greedy(self, window_idx, n_steps):
        """generate.py:24-43 with np.argmax in place of np.random.choice
        (_tests_/faster_generation/generate.py:37); returns (n, n_steps) ints."""
        logits = self.prime(window_idx)
        out = np.empty((self.n, n_steps), dtype=np.int32)
        for s in range(n_steps):
            nxt = np.argmax(logits, axis=1).astype(np.int32)
            out[:, s] = nxt
            if s + 1 < n_steps:
                logits = self.step(nxt)
        return out
