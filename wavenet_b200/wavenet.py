"""Host-side mirror of the reference's wavenet.py object surface, backed by the
sm_100a C ABI (include/wavenet_b200.h).  Names, argument meaning and error
behaviour follow /root/reference/wavenet.py (cited as wavenet.py:line); the
arithmetic runs in libwavenet_b200.so only -- there is no CPU path: every
compute method raises unless to_gpu() was called on a CUDA device.

Shapes seen by the caller are the reference's (B, C, 1, W); on the device the
tensors are channels-last (B, W, C) and `.data` is a zero-copy permuted view.
"""
import ctypes as C
import json
import math
import os

import numpy as np
import torch

from . import _lib
from ._lib import check
from .dist import allreduce_sum_


# --------------------------------------------------------------------------
class Params(object):
    """wavenet.py:100-173 (Python-3 restatement of the attribute bag)."""

    def __init__(self, dict=None):
        self.quantization_steps = 256
        self.sampling_rate = 8000
        self.causal_conv_no_bias = True
        self.causal_conv_filter_width = 2
        self.causal_conv_channels = [128]
        self.residual_conv_dilation_no_bias = True
        self.residual_conv_projection_no_bias = True
        self.residual_conv_filter_width = 2
        self.residual_conv_channels = [32, 32, 32, 32, 32, 32, 32, 32, 32]
        self.residual_num_blocks = 2
        self.softmax_conv_no_bias = False
        self.softmax_conv_channels = [128, 256]
        self.optimizer = "adam"
        self.weight_decay = 0
        self.momentum = 0.9
        self.gradient_clipping = 1.0
        if dict:
            self.from_dict(dict)

    def from_dict(self, dict):
        for attr, value in dict.items():
            if hasattr(self, attr):
                setattr(self, attr, value)

    def to_dict(self):
        return {attr: value for attr, value in self.__dict__.items()}

    def dump(self):
        print("params:")
        for attr, value in self.__dict__.items():
            print("	{}: {}".format(attr, value))

    def check(self):
        base = Params()
        for attr, value in self.__dict__.items():
            if not hasattr(base, attr):
                raise Exception("invalid parameter '{}'".format(attr))
        if self.quantization_steps != self.softmax_conv_channels[-1]:
            raise Exception("quantization_steps != softmax_conv_channels[-1]")

    def to_wn_config(self):
        c = _lib.wn_config()
        c.quantization_steps = int(self.quantization_steps)
        c.n_causal = len(self.causal_conv_channels)
        if c.n_causal > _lib.WN_MAX_CAUSAL or len(self.residual_conv_channels) > _lib.WN_MAX_LAYERS \
                or len(self.softmax_conv_channels) > _lib.WN_MAX_HEAD:
            raise Exception("network has more layers than the C ABI supports")
        for i, v in enumerate(self.causal_conv_channels):
            c.causal_channels[i] = int(v)
        c.causal_filter_width = int(self.causal_conv_filter_width)
        c.causal_no_bias = int(bool(self.causal_conv_no_bias))
        c.n_res_layers = len(self.residual_conv_channels)
        for i, v in enumerate(self.residual_conv_channels):
            c.residual_channels[i] = int(v)
        c.residual_num_blocks = int(self.residual_num_blocks)
        c.residual_filter_width = int(self.residual_conv_filter_width)
        c.residual_dilation_no_bias = int(bool(self.residual_conv_dilation_no_bias))
        c.residual_projection_no_bias = int(bool(self.residual_conv_projection_no_bias))
        c.n_softmax = len(self.softmax_conv_channels)
        for i, v in enumerate(self.softmax_conv_channels):
            c.softmax_channels[i] = int(v)
        c.softmax_no_bias = int(bool(self.softmax_conv_no_bias))
        return c


# --------------------------------------------------------------------------
class Variable(object):
    """Stand-in for chainer.Variable: `.data` holds the array.  `_tag` records which
    tape phase produced it so later phases can continue on the device tape."""

    def __init__(self, data, tag=None):
        self.data = data
        self._tag = tag

    def to_cpu(self):
        if isinstance(self.data, torch.Tensor):
            self.data = self.data.detach().cpu().numpy()

    def __float__(self):
        return float(self.data)


class _Param(object):
    def __init__(self, net, name):
        self._net, self.name = net, name

    @property
    def data(self):
        return self._net._view(self._net._params, self.name)

    @property
    def grad(self):
        return None if self._net._grads is None else self._net._view(self._net._grads, self.name)


class Link(object):
    """One convolution link (wavenet.py:263-342, 439-440, 455): exposes W and b."""

    def __init__(self, net, link_name, has_bias, **attrs):
        self.name = link_name
        self.W = _Param(net, link_name + "/W")
        self.b = _Param(net, link_name + "/b") if has_bias else None
        self.__dict__.update(attrs)


class ResidualConvLayer(object):
    """wavenet.py:344-368 (container of wf, wg, projection_block, projection_softmax)."""
    pass


class Adam(object):
    """Hyper-parameters of chainer.optimizers.Adam as selected at wavenet.py:83,475."""

    def __init__(self, alpha=0.0001, beta1=0.9, beta2=0.999, eps=1e-8):
        self.alpha, self.beta1, self.beta2, self.eps = alpha, beta1, beta2, eps
        self.t = 0

    @property
    def lr(self):
        fix1 = 1.0 - self.beta1 ** self.t
        fix2 = 1.0 - self.beta2 ** self.t
        return self.alpha * math.sqrt(fix2) / fix1


def get_optimizer(name, lr, momentum=0.9):
    """wavenet.py:81-98: only "adam" (the default, wavenet.py:143) is on the hot path."""
    if name.lower() == "adam":
        return Adam(alpha=lr, beta1=momentum)
    raise Exception("optimizer '{}' is outside the B200 hot path (only 'adam')".format(name))


def _ptr(t):
    return C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# --------------------------------------------------------------------------
class WaveNet(object):
    """wavenet.py:370-639."""

    def __init__(self, params, seed=None):
        params.check()
        self.params = params
        self._libh = _lib.load()
        h = C.c_void_p()
        try:
            check(self._libh.wn_create(C.byref(params.to_wn_config()), C.byref(h)))
        except _lib.WaveNetError as e:
            raise Exception(str(e))
        self._h = h
        self._gpu = False
        self._device = None
        self._grads = None
        self._gacc = None
        self._ws = None
        self._ws_key = None
        self._keep = {}
        self._graphs = {}
        self.data_parallel = False
        self.create_network(seed)
        self.setup_optimizer()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._libh.wn_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # ---- layout -------------------------------------------------------------------
    def _layout(self):
        n = self._libh.wn_num_params(self._h)
        arr = (_lib.wn_param_desc * n)()
        check(self._libh.wn_param_layout(self._h, arr, n))
        out = {}
        for d in arr:
            out[d.name.decode()] = (int(d.offset), int(d.numel), tuple(d.shape[i] for i in range(d.ndim)))
        return out

    def _view(self, flat, name):
        off, n, shape = self.layout[name]
        return flat[off:off + n].reshape(shape)

    def create_network(self, seed=None):
        """wavenet.py:379-455: link objects + Chainer-2 default init (LeCunNormal, bias 0)."""
        p = self.params
        self.layout = self._layout()
        self.flat_size = int(self._libh.wn_flat_size(self._h))
        rng = np.random.RandomState(seed) if seed is not None else np.random
        host = np.zeros(self.flat_size, dtype=np.float32)
        for name, (off, n, shape) in self.layout.items():
            if name.endswith("/W"):
                fan_in = shape[1] * shape[2] * shape[3]
                host[off:off + n] = (rng.normal(0, math.sqrt(1.0 / fan_in), size=n)).astype(np.float32)
        self._params = torch.from_numpy(host)
        self._m = None
        self._v = None
        self.causal_conv_layers = []
        for i in range(len(p.causal_conv_channels)):
            self.causal_conv_layers.append(Link(self, "causal_{}".format(i), not p.causal_conv_no_bias,
                                                filter_width=p.causal_conv_filter_width, dilation=1))
        self.residual_blocks = []
        k = p.residual_conv_filter_width
        for j in range(p.residual_num_blocks):
            block = []
            for i in range(len(p.residual_conv_channels)):
                layer = ResidualConvLayer()
                base = "residual_{}_block_{}_".format(j, i)
                layer.wf = Link(self, base + "wf", not p.residual_conv_dilation_no_bias, filter_width=k, dilation=k ** i)
                layer.wg = Link(self, base + "wg", not p.residual_conv_dilation_no_bias, filter_width=k, dilation=k ** i)
                layer.projection_block = Link(self, base + "projection_block", not p.residual_conv_projection_no_bias)
                layer.projection_softmax = Link(self, base + "projection_softmax", not p.residual_conv_projection_no_bias)
                block.append(layer)
            self.residual_blocks.append(block)
        self.softmax_conv_layers = [Link(self, "softmax_{}".format(i), not p.softmax_conv_no_bias)
                                    for i in range(len(p.softmax_conv_channels) - 1)]

    def setup_optimizer(self):
        """wavenet.py:457-481."""
        self.optimizer = get_optimizer(self.params.optimizer, 0.0001, self.params.momentum)

    def update_laerning_rate(self, lr):
        self.optimizer.alpha = lr

    def update_momentum(self, momentum):
        self.optimizer.beta1 = momentum

    # ---- parameters in / out ----------------------------------------------------------
    def set_weights(self, weights):
        """Inject weights given under the reference link names ({'causal_0/W': array, ...})."""
        for name, arr in weights.items():
            if name not in self.layout:
                raise Exception("unknown parameter '{}'".format(name))
            off, n, shape = self.layout[name]
            a = np.asarray(arr, dtype=np.float32)
            if a.size != n:
                raise Exception("shape mismatch for '{}'".format(name))
            self._params[off:off + n] = torch.from_numpy(a.reshape(-1).copy()).to(self._params.device)

    def get_weights(self):
        flat = self._params.detach().cpu().numpy()
        return {name: flat[off:off + n].reshape(shape).copy() for name, (off, n, shape) in self.layout.items()}

    def get_grads(self):
        flat = self._grads.detach().cpu().numpy()
        return {name: flat[off:off + n].reshape(shape).copy() for name, (off, n, shape) in self.layout.items()}

    # ---- device ------------------------------------------------------------------------
    def to_gpu(self, device=None):
        """wavenet.py:521-523.  Allocates the flat parameter / gradient / Adam buffers."""
        if not torch.cuda.is_available():
            raise Exception("wavenet_b200 has no CPU path: a CUDA device (sm_100a) is required")
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        # the C ABI launches on the CURRENT device's stream (like chain.to_gpu() after cuda.get_device(n).use(),
        # train_audio/model.py:57-59): make the selected device current and re-create the handle there so that
        # sm_count and every launch belong to it
        torch.cuda.set_device(dev)
        if self._device is not None and self._device != dev:
            self._ws, self._ws_key = None, None
        self._graphs = {}
        self._device = dev
        check(self._libh.wn_set_device_info(self._h))
        self._params = self._params.to(dev)
        self._grads = torch.zeros_like(self._params)
        self._m = torch.zeros_like(self._params) if self._m is None else self._m.to(dev)
        self._v = torch.zeros_like(self._params) if self._v is None else self._v.to(dev)
        self._scratch = torch.zeros(int(self._libh.wn_optim_scratch_bytes(self._h)), dtype=torch.uint8, device=dev)
        self._loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self._gacc = None
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self._gpu = True

    @property
    def gpu_enabled(self):
        return bool(torch.cuda.is_available() and self._gpu)

    def set_precision(self, name):
        """'fp32' (SIMT FFMA, exact), 'fp16x2' (tcgen05 on split fp16 operands: fp32-grade, meets the 1e-4 logit /
        1e-3 gradient gates) or 'tf32' (single-pass tcgen05, 1e-2 logits)."""
        check(self._libh.wn_set_precision(self._h, {"fp32": _lib.WN_PREC_FP32, "tf32": _lib.WN_PREC_TF32,
                                                    "fp16x2": _lib.WN_PREC_F16X2}[name]))
        self._graphs = {}          # a captured step replays the kernels of the precision it was recorded under

    def set_deterministic(self, on=True):
        """Bit-reproducible training (gradients, weights) run to run: fixed-order reductions instead of atomics
        (wn_set_deterministic; fused fp16x2 shape only, ~0.4 ms per step and ~150 x the gradient buffer of scratch)."""
        self._need_gpu()
        if on:
            nbytes = int(self._libh.wn_det_scratch_bytes(self._h))
            self._det_scratch = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            check(self._libh.wn_set_deterministic(self._h, 1, _ptr(self._det_scratch)))
        else:
            check(self._libh.wn_set_deterministic(self._h, 0, None))
            self._det_scratch = None
        self._graphs = {}          # captured steps hold the launch sequence of the mode they were recorded under

    def _need_gpu(self):
        if not self.gpu_enabled:
            raise Exception("wavenet_b200 has no CPU path: call to_gpu() on a CUDA device first")

    def _bind(self, B, W):
        if self._ws_key == (B, W):
            return
        nbytes = int(self._libh.wn_workspace_bytes(self._h, B, W))
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
        check(self._libh.wn_bind_workspace(self._h, _ptr(self._ws), nbytes, B, W))
        self._graphs = {}          # captured steps hold TMA descriptors / offsets of the previous binding
        self._ws_key = (B, W)
        self._B, self._W = B, W

    # ---- input handling -------------------------------------------------------------
    def to_variable(self, x):
        """wavenet.py:537-542."""
        if not isinstance(x, Variable):
            x = Variable(x)
        return x

    def to_numpy(self, x):
        """wavenet.py:544-549."""
        if isinstance(x, Variable):
            x = x.data
        if isinstance(x, torch.Tensor):
            x = x.detach().cpu().numpy()
        return x

    def get_batchsize(self, x):
        if isinstance(x, Variable):
            return x.data.shape[0]
        return x.shape[0]

    def _to_indices(self, x):
        """Accepts (B, W) integer samples or the reference's (B, Q, 1, W) one-hot image
        (data.py:61-68) and returns int32 indices on the device."""
        if isinstance(x, Variable):
            x = x.data
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        if x.dim() == 2:
            return x.to(self._device, dtype=torch.int32).contiguous()
        if x.dim() != 4 or x.shape[1] != self.params.quantization_steps or x.shape[2] != 1:
            raise Exception("expected (B, W) samples or a (B, {}, 1, W) one-hot image".format(self.params.quantization_steps))
        oh = x.to(self._device, dtype=torch.float32).contiguous()
        B, Q, _, W = oh.shape
        idx = torch.empty((B, W), dtype=torch.int32, device=self._device)
        check(self._libh.wn_onehot_to_index(_ptr(oh), B, Q, W, _ptr(idx), _stream()))
        return idx

    @staticmethod
    def _as_bc1w(t):
        return t.permute(0, 2, 1).unsqueeze(2)      # (B, W, C) -> (B, C, 1, W) view

    def _to_bwc(self, x):
        if isinstance(x, Variable):
            x = x.data
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        x = x.to(self._device, dtype=torch.float32)
        return x[:, :, 0, :].permute(0, 2, 1).contiguous()

    # ---- forward (wavenet.py:556-593) -------------------------------------------------
    def forward_one_step(self, x_batch, apply_softmax=True, as_numpy=False):
        causal_output = self.forward_causal_block(x_batch)
        residual_output, sum_skip_connections = self.forward_residual_block(causal_output)
        softmax_output = self.forward_softmax_block(sum_skip_connections, apply_softmax=apply_softmax)
        if as_numpy:
            return self.to_numpy(softmax_output)
        return softmax_output

    def forward_causal_block(self, x_batch):
        self._need_gpu()
        idx = self._to_indices(x_batch)
        B, W = idx.shape
        self._bind(B, W)
        self._keep["idx"] = idx
        out = torch.empty((B, W, self.params.causal_conv_channels[-1]), dtype=torch.float32, device=self._device)
        check(self._libh.wn_forward_causal_block(self._h, _ptr(self._params), _ptr(idx), _ptr(out), _stream()))
        # the tag names the tape contents this Variable stands for: a later pass of the same shape replaces them, and a
        # Variable with a stale generation then goes through the external-input path (its .data is a real copy)
        self._tape_gen = getattr(self, "_tape_gen", 0) + 1
        return Variable(self._as_bc1w(out), tag=("causal", B, W, self._tape_gen))

    def forward_residual_block(self, x_batch):
        self._need_gpu()
        x_batch = self.to_variable(x_batch)
        R, S = self.params.causal_conv_channels[-1], self.params.softmax_conv_channels[0]
        tag = x_batch._tag
        if (tag is not None and tag[0] == "causal" and self._ws_key == tag[1:3]
                and tag[3] == getattr(self, "_tape_gen", 0)):
            B, W = tag[1:3]
            xin = None
        else:
            xin = self._to_bwc(x_batch)
            B, W = xin.shape[0], xin.shape[1]
            self._bind(B, W)
            self._tape_gen = getattr(self, "_tape_gen", 0) + 1
        out = torch.empty((B, W, R), dtype=torch.float32, device=self._device)
        skip = torch.empty((B, W, S), dtype=torch.float32, device=self._device)
        check(self._libh.wn_forward_residual_block(self._h, _ptr(self._params), _ptr(xin), _ptr(out), _ptr(skip), _stream()))
        gen = self._tape_gen
        return (Variable(self._as_bc1w(out), tag=("out", B, W, W, gen)),
                Variable(self._as_bc1w(skip), tag=("skip", B, W, W, gen)))

    def forward_softmax_block(self, x_batch, apply_softmax=True):
        self._need_gpu()
        x_batch = self.to_variable(x_batch)
        Q = self.params.quantization_steps
        tag = x_batch._tag
        if (tag is not None and tag[0] == "skip" and self._ws_key == tag[1:3]
                and tag[4] == getattr(self, "_tape_gen", 0)):
            B, W, T = tag[1], tag[2], tag[3]
            xin = None
        else:
            xin = self._to_bwc(x_batch)
            B, T = xin.shape[0], xin.shape[1]
            if self._ws_key is None or self._ws_key[0] != B or self._ws_key[1] < T:
                self._bind(B, T)
        out = torch.empty((B, T, Q), dtype=torch.float32, device=self._device)
        check(self._libh.wn_forward_softmax_block(self._h, _ptr(self._params), _ptr(xin), T, int(bool(apply_softmax)),
                                                  _ptr(out), _stream()))
        return Variable(self._as_bc1w(out), tag=("logits", B, T, bool(apply_softmax)))

    # ---- pad / slice (wavenet.py:531-535) ------------------------------------------------
    def slice_1d(self, x, cut=0):
        if cut < 1:
            raise Exception("CausalSlice1d: cut cannot be less than one.")
        x = self.to_variable(x)
        tag = x._tag
        if tag is not None and tag[0] in ("skip", "out"):
            tag = (tag[0], tag[1], tag[2], tag[3] - cut, tag[4])
        else:
            tag = None
        return Variable(x.data[:, :, :, cut:], tag=tag)

    def padding_1d(self, x, pad=0):
        x = self.to_variable(x)
        d = x.data
        if isinstance(d, np.ndarray):
            out = np.zeros(d.shape[:3] + (d.shape[3] + pad,), dtype=np.float32)
            out[:, :, :, pad:] = d
        else:
            out = torch.zeros(tuple(d.shape[:3]) + (d.shape[3] + pad,), dtype=torch.float32, device=d.device)
            out[:, :, :, pad:] = d
        return Variable(out)

    # ---- loss / update (wavenet.py:597-617, 515-519) -----------------------------------------
    def cross_entropy(self, raw_network_output, target_signal_data):
        if isinstance(target_signal_data, Variable):
            raise Exception("target_signal_data cannot be Variable")
        self._need_gpu()
        raw_network_output = self.to_variable(raw_network_output)
        target_width = target_signal_data.shape[1]
        if raw_network_output.data.shape[3] != target_width:
            raise Exception("raw_network_output.width != target.width")
        tag = raw_network_output._tag
        if tag is None or tag[0] != "logits" or tag[3]:
            raise Exception("cross_entropy expects the raw output of forward_softmax_block(apply_softmax=False)")
        if isinstance(target_signal_data, np.ndarray):
            target_signal_data = torch.from_numpy(np.ascontiguousarray(target_signal_data))
        tgt = target_signal_data.to(self._device, dtype=torch.int32).contiguous()
        self._keep["tgt"] = tgt
        check(self._libh.wn_cross_entropy(self._h, _ptr(tgt), _ptr(self._loss), _stream()))
        return Variable(self._loss[0], tag=("loss",))

    def backward(self):
        """loss.backward(): fills the flat gradient buffer (no update)."""
        check(self._libh.wn_backward(self._h, _ptr(self._params), _ptr(self._grads), _stream()))

    def update(self, grads=None, micro_batches=1):
        """Hooks + Adam (wavenet.py:477-480, Chainer GradientMethod.update).  `grads`: another flat gradient buffer (the
        accumulator of train_step_accumulated) holding the SUM over `micro_batches` micro-batch gradients."""
        p, opt = self.params, self.optimizer
        grad_scale = 1.0
        if grads is None:
            grads = self._grads
        if self.data_parallel:
            world = int(self._libh.wn_comm_world(self._h))
            if world > 1:        # communicator behind the C ABI (dist.init_comm): exchange + hooks + Adam in one entry point
                opt.t += 1       # (NCCL all-reduce, or the fused one-shot peer-memory all-reduce when it is enabled)
                check(self._libh.wn_allreduce_clip_adam_step(
                    self._h, _ptr(self._params), _ptr(grads), _ptr(self._m), _ptr(self._v), opt.t, opt.alpha, opt.beta1,
                    opt.beta2, opt.eps, float(p.weight_decay), float(p.gradient_clipping), 1.0 / (world * micro_batches),
                    _ptr(self._scratch), _ptr(self._norm), _stream()))
                return
            grad_scale = allreduce_sum_(grads)      # no communicator in the library: torch.distributed (host-side tests)
        grad_scale /= micro_batches
        opt.t += 1
        check(self._libh.wn_clip_adam_step(self._h, _ptr(self._params), _ptr(grads), _ptr(self._m), _ptr(self._v),
                                           opt.t, opt.alpha, opt.beta1, opt.beta2, opt.eps, float(p.weight_decay),
                                           float(p.gradient_clipping), grad_scale, _ptr(self._scratch),
                                           _ptr(self._norm), _stream()))

    def backprop(self, loss):
        if not isinstance(loss, Variable):
            loss = loss()
        self.backward()
        self.update()

    def create_batch(self, signal, starts, input_width, target_width):
        """train_audio/train.py:14-22 on the device.  `signal`: int32 cuda tensor (the silence-padded quantised
        clip, uploaded once per file); `starts`: the host-drawn crop indices (np.random.randint as in the
        reference).  Returns (input_batch (B, iw+tw), target_batch (B, tw)) int32 cuda tensors."""
        self._need_gpu()
        st = torch.as_tensor(np.asarray(starts, dtype=np.int32)).to(self._device, non_blocking=True)
        B = int(st.shape[0])
        x = torch.empty((B, input_width + target_width), dtype=torch.int32, device=self._device)
        t = torch.empty((B, target_width), dtype=torch.int32, device=self._device)
        check(self._libh.wn_crop_batch(_ptr(signal), int(signal.numel()), _ptr(st), B, int(input_width), int(target_width),
                                       _ptr(x), _ptr(t), _stream()))
        return x, t

    def _fwd_bwd(self, x_idx, target, T):
        self._tape_gen = getattr(self, "_tape_gen", 0) + 1      # the fused step rewrites the tape: older Variables are stale
        check(self._libh.wn_forward_loss(self._h, _ptr(self._params), _ptr(x_idx), _ptr(target), T, _ptr(self._loss),
                                         None, _stream()))
        self.backward()

    def train_step_accumulated(self, micro_batches, train_width=None):
        """One optimiser step over a global batch that is processed as several micro-batches [(x_idx, target), ...] of the
        same shape (BASELINE config 5 on fewer than 8 GPUs: 256 x 16000 does not fit one tape): every micro-batch runs the
        captured forward + loss + backward, its gradient is added to an accumulator (wn_accumulate_grads), and ONE
        all-reduce + clip + Adam follows on the mean gradient.  Returns the mean loss (device tensor)."""
        if self._gacc is None or self._gacc.shape != self._grads.shape:
            self._gacc = torch.zeros_like(self._grads)
            self._lacc = torch.zeros_like(self._loss)
        self._gacc.zero_()
        self._lacc.zero_()
        for x_idx, target in micro_batches:
            self.train_step(x_idx, target, train_width, _update=False)
            check(self._libh.wn_accumulate_grads(self._h, _ptr(self._grads), _ptr(self._gacc), _stream()))
            self._lacc += self._loss
        self.update(self._gacc, len(micro_batches))
        return self._lacc / len(micro_batches)

    def train_step(self, x_idx, target, train_width=None, _update=True):
        """One fused train.py:58-80 step on int32 device tensors; returns the loss tensor (no sync).

        The ~200 launches of forward + loss + backward are captured once per (B, W, T) shape into a CUDA graph
        and replayed (launch-bound shapes such as the reference's 16 x 757 batches gain most); the optimiser
        hooks + Adam stay outside because their step size changes every call.  Set `use_cuda_graph = False`
        to launch eagerly."""
        self._need_gpu()
        B, W = x_idx.shape
        T = W if train_width is None else train_width
        if x_idx.dtype != torch.int32 or target.dtype != torch.int32:
            raise Exception("train_step expects int32 device tensors")
        if tuple(target.shape) != (B, T):
            raise Exception("raw_network_output.width != target.width")
        if torch.cuda.current_device() != self._device.index:
            torch.cuda.set_device(self._device)
        self._bind(B, W)
        self._tape_gen = getattr(self, "_tape_gen", 0) + 1      # the step (eager or replayed) rewrites the tape
        key = (B, W, T, int(self._libh.wn_get_precision(self._h)))
        if not getattr(self, "use_cuda_graph", True):
            self._keep["idx"], self._keep["tgt"] = x_idx, target
            self._fwd_bwd(x_idx, target, T)
        else:
            g = self._graphs.get(key) if hasattr(self, "_graphs") else None
            if not hasattr(self, "_graphs"):
                self._graphs = {}
            if g is None:
                # static input buffers + one eager warm-up (one-time uploads / attribute calls), then capture
                sx, st = torch.empty_like(x_idx), torch.empty_like(target)
                sx.copy_(x_idx)
                st.copy_(target)
                self._fwd_bwd(sx, st, T)
                graph = None
                try:
                    torch.cuda.synchronize()
                    graph = torch.cuda.CUDAGraph()
                    n0 = int(self._libh.wn_launch_count(0))
                    with torch.cuda.graph(graph):
                        self._fwd_bwd(sx, st, T)
                    self._graph_launches = int(self._libh.wn_launch_count(0)) - n0   # kernels recorded in the graph
                    self._libh.wn_launch_count_add(-self._graph_launches)            # capture itself launched nothing
                except Exception:
                    graph = None                      # capture unsupported here: stay eager
                    torch.cuda.synchronize()
                self._graphs = {key: (graph, sx, st)}     # one shape at a time (the tape is re-bound on shape change)
                if graph is None:
                    if _update:
                        self.update()
                    return self._loss
                g = self._graphs[key]
            graph, sx, st = g
            if graph is None:
                self._fwd_bwd(x_idx, target, T)
            else:
                sx.copy_(x_idx, non_blocking=True)
                st.copy_(target, non_blocking=True)
                graph.replay()
                self._libh.wn_launch_count_add(self._graph_launches)
        if _update:
            self.update()
        return self._loss

    # ---- checkpoint (wavenet.py:619-639) --------------------------------------------------------------------------
    # Stored as .npz under the reference link names; when h5py is importable the reference's own files are written and
    # read as well: `wavenet.model` / `wavenet.opt` in the layout chainer.serializers.save_hdf5 produces for the chain
    # registered at wavenet.py:461-472 (group "<link name>", datasets "W" / "b"; optimizer: "t", "epoch" and per-parameter
    # groups "<link name>/W" with the Adam moments "m", "v"), so weights trained with the Chainer reference load here.
    @staticmethod
    def _h5py():
        try:
            import h5py
            return h5py
        except Exception:
            return None

    def save(self, model_dir="./"):
        try:
            os.mkdir(model_dir)
        except Exception:
            pass
        weights = self.get_weights()
        np.savez(model_dir + "/wavenet.model.npz", **weights)
        opt = {"t": np.int64(self.optimizer.t)}
        moments = {}
        if self._m is not None:
            m, v = self._m.detach().cpu().numpy(), self._v.detach().cpu().numpy()
            for name, (off, n, shape) in self.layout.items():
                moments[name] = (m[off:off + n].reshape(shape), v[off:off + n].reshape(shape))
                opt[name + "/m"], opt[name + "/v"] = moments[name]
        np.savez(model_dir + "/wavenet.opt.npz", **opt)
        h5py = self._h5py()
        if h5py is not None:
            with h5py.File(model_dir + "/wavenet.model", "w") as f:
                for name, arr in weights.items():            # "causal_0/W" -> group causal_0, dataset W
                    f.create_dataset(name, data=arr)
            with h5py.File(model_dir + "/wavenet.opt", "w") as f:
                f.create_dataset("t", data=np.int32(self.optimizer.t))
                f.create_dataset("epoch", data=np.int32(0))
                for name, (mm, vv) in moments.items():
                    f.create_dataset(name + "/m", data=mm)
                    f.create_dataset(name + "/v", data=vv)

    def _set_moments(self, get):
        dev = self._params.device
        m = torch.zeros(self.flat_size, dtype=torch.float32)
        v = torch.zeros(self.flat_size, dtype=torch.float32)
        for name, (off, n, shape) in self.layout.items():
            mv = get(name)
            if mv is not None:
                m[off:off + n] = torch.from_numpy(np.asarray(mv[0], dtype=np.float32).reshape(-1))
                v[off:off + n] = torch.from_numpy(np.asarray(mv[1], dtype=np.float32).reshape(-1))
        self._m, self._v = m.to(dev), v.to(dev)

    def load(self, model_dir="./"):
        h5py = self._h5py()
        filename = model_dir + "/wavenet.model.npz"
        if os.path.isfile(filename):
            print("loading", filename, "...")
            with np.load(filename) as f:
                self.set_weights({k: f[k] for k in f.files})
        elif h5py is not None and os.path.isfile(model_dir + "/wavenet.model"):
            print("loading", model_dir + "/wavenet.model", "...")
            with h5py.File(model_dir + "/wavenet.model", "r") as f:
                self.set_weights({name: np.asarray(f[name]) for name in self.layout if name in f})
        filename = model_dir + "/wavenet.opt.npz"
        if os.path.isfile(filename):
            print("loading", filename, "...")
            with np.load(filename) as f:
                self.optimizer.t = int(f["t"])
                self._set_moments(lambda name: (f[name + "/m"], f[name + "/v"]) if name + "/m" in f.files else None)
        elif h5py is not None and os.path.isfile(model_dir + "/wavenet.opt"):
            print("loading", model_dir + "/wavenet.opt", "...")
            with h5py.File(model_dir + "/wavenet.opt", "r") as f:
                self.optimizer.t = int(np.asarray(f["t"]))
                self._set_moments(lambda name: (np.asarray(f[name + "/m"]), np.asarray(f[name + "/v"]))
                                  if name + "/m" in f else None)
