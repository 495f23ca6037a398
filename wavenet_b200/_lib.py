"""ctypes binding of libwavenet_b200.so (include/wavenet_b200.h).

The library is the product: there is no CPU fallback.  If it is missing the
import fails loudly with the build command.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WN_LIB_PATH") or os.path.join(_HERE, "libwavenet_b200.so")   # override: A/B builds

WN_MAX_CAUSAL, WN_MAX_LAYERS, WN_MAX_HEAD, WN_NAME_LEN = 8, 32, 8, 64
WN_PREC_FP32, WN_PREC_TF32, WN_PREC_F16X2 = 0, 1, 2
WN_GEN_GREEDY, WN_GEN_SAMPLE = 0, 1


class wn_config(C.Structure):
    _fields_ = [
        ("quantization_steps", C.c_int32),
        ("n_causal", C.c_int32),
        ("causal_channels", C.c_int32 * WN_MAX_CAUSAL),
        ("causal_filter_width", C.c_int32),
        ("causal_no_bias", C.c_int32),
        ("n_res_layers", C.c_int32),
        ("residual_channels", C.c_int32 * WN_MAX_LAYERS),
        ("residual_num_blocks", C.c_int32),
        ("residual_filter_width", C.c_int32),
        ("residual_dilation_no_bias", C.c_int32),
        ("residual_projection_no_bias", C.c_int32),
        ("n_softmax", C.c_int32),
        ("softmax_channels", C.c_int32 * WN_MAX_HEAD),
        ("softmax_no_bias", C.c_int32),
    ]


class wn_param_desc(C.Structure):
    _fields_ = [
        ("name", C.c_char * WN_NAME_LEN),
        ("offset", C.c_int64),
        ("numel", C.c_int64),
        ("ndim", C.c_int32),
        ("shape", C.c_int32 * 4),
    ]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_L = C.c_int64

# name -> (restype, argtypes): every symbol include/wavenet_b200.h declares
SIGNATURES = {
    "wn_last_error": (C.c_char_p, []),
    "wn_version": (_I, []),
    "wn_launch_count": (_L, [_I]),
    "wn_launch_count_add": (_L, [_L]),
    "wn_create": (_I, [C.POINTER(wn_config), C.POINTER(_P)]),
    "wn_destroy": (_I, [_P]),
    "wn_set_device_info": (_I, [_P]),
    "wn_set_precision": (_I, [_P, _I]),
    "wn_get_precision": (_I, [_P]),
    "wn_tc_active": (_I, [_P]),
    "wn_flat_size": (_L, [_P]),
    "wn_param_elems": (_L, [_P]),
    "wn_num_params": (_I, [_P]),
    "wn_param_layout": (_I, [_P, C.POINTER(wn_param_desc), _I]),
    "wn_receptive_width": (_I, [_P]),
    "wn_input_width": (_I, [_P]),
    "wn_zero_prefix": (_I, [_I, _I, _I]),
    "wn_workspace_bytes": (_L, [_P, _I, _I]),
    "wn_bind_workspace": (_I, [_P, _P, _L, _I, _I]),
    "wn_forward_causal_block": (_I, [_P, _P, _P, _P, _P]),
    "wn_forward_residual_block": (_I, [_P, _P, _P, _P, _P, _P]),
    "wn_forward_softmax_block": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "wn_cross_entropy": (_I, [_P, _P, _P, _P]),
    "wn_backward": (_I, [_P, _P, _P, _P]),
    "wn_forward_loss": (_I, [_P, _P, _P, _P, _I, _P, _P, _P]),
    "wn_tc_layer_forward": (_I, [_P, _I, _P]),
    "wn_tc_skip_gemm": (_I, [_P, _P]),
    "wn_tc_gate_backward_layer": (_I, [_P, _I, _P, _P]),
    "wn_tcs_layer_forward": (_I, [_P, _I, _P]),
    "wn_tcs_skip_gemm": (_I, [_P, _P]),
    "wn_tcs_debug_gemm": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P, _I, _P, _I, _P]),
    "wn_tcs_debug_wgrad": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _F, _P, _P]),
    "wn_comm_available": (_I, []),
    "wn_comm_unique_id": (_I, [C.c_char_p]),
    "wn_comm_init": (_I, [_P, C.c_char_p, _I, _I]),
    "wn_comm_world": (_I, [_P]),
    "wn_allreduce_grads": (_I, [_P, _P, _P]),
    "wn_accumulate_grads": (_I, [_P, _P, _P, _P]),
    "wn_comm_peer_enabled": (_I, [_P]),
    "wn_allreduce_clip_adam_step": (_I, [_P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _F, _F, _P, _P, _P]),
    "wn_det_scratch_bytes": (_L, [_P]),
    "wn_set_deterministic": (_I, [_P, _I, _P]),
    "wn_comm_destroy": (_I, [_P]),
    "wn_optim_scratch_bytes": (_L, [_P]),
    "wn_clip_adam_step": (_I, [_P, _P, _P, _P, _P, _I, _F, _F, _F, _F, _F, _F, _F, _P, _P, _P]),
    "wn_gen_create": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "wn_gen_destroy": (_I, [_P]),
    "wn_gen_state_bytes": (_L, [_P]),
    "wn_gen_bind_state": (_I, [_P, _P, _L]),
    "wn_gen_prime": (_I, [_P, _P, _P, _P, _P]),
    "wn_gen_prime_part": (_I, [_P, _P, _P, _I, _I, _P, _P]),
    "wn_gen_mma_capacity": (_I, [_I]),
    "wn_gen_step": (_I, [_P, _P, _P, _I, _P, _P]),
    "wn_gen_logits": (_I, [_P, _P, _P]),
    "wn_gen_run": (_I, [_P, _P, _I, _I, C.c_uint64, _P, _P]),
    "wn_crop_batch": (_I, [_P, _L, _P, _I, _I, _I, _P, _P, _P]),
    "wn_onehot_to_index": (_I, [_P, _I, _I, _I, _P, _P]),
    "wn_mulaw_encode": (_I, [_P, _L, _I, _P, _P]),
    "wn_mulaw_decode": (_I, [_P, _L, _I, C.c_double, _P, _P]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libwavenet_b200.so is not built (no CPU fallback exists). Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C wavenet_b200/csrc`.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class WaveNetError(Exception):
    """Raised for any non-zero return of the C ABI (the reference raises bare Exception)."""

    def __init__(self, code, msg):
        Exception.__init__(self, msg)
        self.code = code


def check(code):
    if code < 0:
        raise WaveNetError(code, load().wn_last_error().decode("utf-8", "replace"))
    return code
