"""ctypes binding of the C ABI in include/iins_b200.h.

The product path loads exactly one library: the nvcc-built, in-tree ``libiins_b200.so`` (sm_100a).
If it is missing it is built with nvcc; if that is impossible the import raises -- there is no CPU
or PyTorch fallback anywhere in this package.
"""
import ctypes as C
import os

ABI_VERSION = 3
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libiins_b200.so")

EXPORTS = [
    "iins_abi_version", "iins_last_error", "iins_validate_config",
    "iins_encoder_num_params", "iins_encoder_ws_floats", "iins_encoder_scratch_floats",
    "iins_encoder_forward", "iins_encoder_backward",
    "iins_decoder_num_params", "iins_decoder_ws_floats", "iins_decoder_scratch_floats",
    "iins_decoder_forward", "iins_decoder_backward",
    "iins_restorer_num_params", "iins_restorer_ws_floats", "iins_restorer_scratch_floats",
    "iins_restorer_forward", "iins_restorer_backward",
    "iins_classifier_num_params", "iins_classifier_ws_floats", "iins_classifier_scratch_floats",
    "iins_classifier_forward", "iins_classifier_backward",
    "iins_loss_forward_backward", "iins_adam_step",
    "iins_adaptive_pool_forward", "iins_adaptive_pool_backward", "iins_accumulate2", "iins_set_stream_concurrency",
    "iins_launch_count", "iins_profile_begin", "iins_profile_collect",
    "iins_set_compute_mode", "iins_get_compute_mode", "iins_profile_shapes", "iins_profile_bytes",
    "iins_encoder2d_ws_floats", "iins_encoder2d_scratch_floats", "iins_encoder2d_forward", "iins_encoder2d_backward",
    "iins_decoder2d_ws_floats", "iins_decoder2d_scratch_floats", "iins_decoder2d_forward", "iins_decoder2d_backward",
    "iins_restorer_conv_ws_floats", "iins_restorer_conv_scratch_floats", "iins_restorer_conv_forward", "iins_restorer_conv_backward",
    "iins_classifier_conv_ws_floats", "iins_classifier_conv_scratch_floats", "iins_classifier_conv_forward",
    "iins_classifier_conv_backward",
    "iins_ctx_create", "iins_ctx_destroy", "iins_ctx_make_current", "iins_ctx_get_current", "iins_ctx_set_compute_mode",
    "iins_ctx_get_compute_mode", "iins_ctx_set_stream_concurrency",
    "iins_set_deferred_join", "iins_join_helpers",
    "iins_restorer_soft_ws_floats", "iins_restorer_soft_scratch_floats", "iins_restorer_soft_forward", "iins_restorer_soft_backward",
]


class IinsConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("batch", "cir_len", "dim", "n_residual", "n_downsample", "env_dim",
                                       "range_dim", "num_classes", "cls_filters", "conv_type")]


class IinsHeadState(C.Structure):
    """iins_head_state (include/iins_b200.h): dropout source + BatchNorm buffers of a Conv1d head."""
    _fields_ = [("training", C.c_int), ("mask1", C.c_void_p), ("mask2", C.c_void_p), ("seed", C.c_uint64), ("offset", C.c_uint64),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("bn_stats", C.c_void_p), ("phase", C.c_int), ("count_scale", C.c_double), ("sample_offset", C.c_int64),
                ("offset_dev", C.c_void_p)]


class IinsError(RuntimeError):
    pass


_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)
_CFG = C.POINTER(IinsConfig)


class IinsLib:
    """Typed view of one loaded shared object exporting the iins_* symbols."""

    def __init__(self, path: str):
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.iins_abi_version.restype = C.c_int
        d.iins_join_helpers.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        d.iins_ctx_create.restype = C.c_void_p
        d.iins_ctx_get_current.restype = C.c_void_p
        d.iins_ctx_destroy.argtypes = [C.c_void_p]
        d.iins_ctx_destroy.restype = None
        d.iins_ctx_make_current.argtypes = [C.c_void_p]
        d.iins_ctx_set_compute_mode.argtypes = [C.c_void_p, C.c_int]
        d.iins_ctx_get_compute_mode.argtypes = [C.c_void_p]
        d.iins_ctx_set_stream_concurrency.argtypes = [C.c_void_p, C.c_int]
        d.iins_last_error.restype = C.c_char_p
        d.iins_validate_config.argtypes = [_CFG]
        for mod in ("encoder", "decoder", "restorer", "classifier"):
            getattr(d, f"iins_{mod}_num_params").argtypes = [_CFG]
            for q in ("ws", "scratch"):
                f = getattr(d, f"iins_{mod}_{q}_floats")
                f.argtypes = [_CFG]
                f.restype = C.c_size_t
        d.iins_encoder_forward.argtypes = [_CFG, _PP, _P, _P, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P]
        d.iins_encoder_backward.argtypes = [_CFG, _PP, _P, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _PP, _P, _P]
        d.iins_decoder_forward.argtypes = [_CFG, _PP, _P, _P, _P, _P, _P]
        d.iins_decoder_backward.argtypes = [_CFG, _PP, _P, _P, _P, _P, _PP, _P, _P, C.c_int, _P, _P]
        if hasattr(d, "iins_encoder2d_forward"):         # (the logic-simulator build of the tests has no 2-D plans)
            for mod in ("encoder2d", "decoder2d"):
                for suffix in ("ws_floats", "scratch_floats"):
                    f = getattr(d, f"iins_{mod}_{suffix}")
                    f.argtypes = [_CFG]
                    f.restype = C.c_size_t
            d.iins_encoder2d_forward.argtypes = d.iins_encoder_forward.argtypes
            d.iins_encoder2d_backward.argtypes = d.iins_encoder_backward.argtypes
            d.iins_decoder2d_forward.argtypes = d.iins_decoder_forward.argtypes
            d.iins_decoder2d_backward.argtypes = d.iins_decoder_backward.argtypes
        d.iins_restorer_forward.argtypes = [_CFG, _PP, _P, _P, _P, _P]
        d.iins_restorer_backward.argtypes = [_CFG, _PP, _P, _P, _P, _PP, _P, C.c_int, _P, _P]
        d.iins_classifier_forward.argtypes = [_CFG, _PP, _P, _P, _P, _P]
        d.iins_classifier_backward.argtypes = [_CFG, _PP, _P, _P, _P, _PP, _P, C.c_int, _P, _P]
        _HS = C.POINTER(IinsHeadState)
        for mod in ("restorer", "classifier"):
            for q in ("ws", "scratch"):
                f = getattr(d, f"iins_{mod}_conv_{q}_floats")
                f.argtypes = [_CFG]
                f.restype = C.c_size_t
            getattr(d, f"iins_{mod}_conv_forward").argtypes = [_CFG, _PP, _P, _P, _P, _HS, _P]
            getattr(d, f"iins_{mod}_conv_backward").argtypes = [_CFG, _PP, _P, _P, _P, _PP, _P, C.c_int, _P, _HS, _P]
        for q in ("ws", "scratch"):
            f = getattr(d, f"iins_restorer_soft_{q}_floats")
            f.argtypes = [_CFG]
            f.restype = C.c_size_t
        d.iins_restorer_soft_forward.argtypes = [_CFG, _PP, _P, _P, _P, _P, _P]
        d.iins_restorer_soft_backward.argtypes = [_CFG, _PP, _P, _P, _P, _P, _PP, _P, C.c_int, _P, _P]
        d.iins_loss_forward_backward.argtypes = [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.c_int,
                                                 C.c_float, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P]
        d.iins_adam_step.argtypes = [_P, _P, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                     C.c_int, _P, _P, C.c_double, C.c_double, C.c_float, C.c_float, C.c_int, _P]
        d.iins_adaptive_pool_forward.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, _P]
        d.iins_adaptive_pool_backward.argtypes = [_P, _P, C.c_int, C.c_int, C.c_int, _P]
        d.iins_accumulate2.argtypes = [_P, _P, C.c_size_t, _P, _P, C.c_size_t, _P]
        d.iins_launch_count.restype = C.c_ulonglong
        d.iins_profile_collect.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int]
        d.iins_profile_bytes.argtypes = [C.POINTER(C.c_double), C.c_int]

    def profile(self, fn):
        """Run fn() with a CUDA-event pair around every kernel launch; returns [(kernel name, ms, flops), ...]."""
        self.check(self.dll.iins_profile_begin(), "profile_begin")
        fn()
        cap = 4096
        names = (C.c_char_p * cap)()
        ms = (C.c_float * cap)()
        fl = (C.c_double * cap)()
        n = self.dll.iins_profile_collect(names, ms, fl, cap)
        self.last_shapes = (C.c_int * (3 * cap))()
        self.dll.iins_profile_shapes(self.last_shapes, cap)
        self.last_bytes = (C.c_double * cap)()
        self.dll.iins_profile_bytes(self.last_bytes, cap)
        return [(names[i].decode(), float(ms[i]), float(fl[i])) for i in range(n)]

    def check(self, rc: int, what: str):
        if rc != 0:
            raise IinsError(f"{what} failed ({rc}): {self.dll.iins_last_error().decode()}")

    def __getattr__(self, name):
        return getattr(self.dll, name)


def ptr(t):
    """Device (or, in the simulator tests, host) address of a tensor, or NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


_lib = None


def get_lib() -> IinsLib:
    """The sm_100a library.  Built in-tree on first use when nvcc is present; otherwise raises."""
    global _lib
    if _lib is None:
        from . import build as _build
        if _build.needs_build():
            _build.build()
        _lib = IinsLib(LIB_PATH)
        if _lib.iins_abi_version() != ABI_VERSION:
            raise IinsError("libiins_b200.so ABI version mismatch")
    return _lib
