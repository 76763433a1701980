"""ctypes binding of include/l2s_vocoder.h (the drop-in C ABI) and include/l2s_hand_off.h (batched file I/O).

The library is never built or searched for implicitly at import time, and
there is no fallback: a missing .so raises ``ImportError`` with the build command.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libl2s_vocoder.so")

L2S_MAX_UPS, L2S_MAX_RK, L2S_MAX_DIL = 8, 4, 4

OK, ERR_INVALID, ERR_SHAPE, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_INDEX = range(8)
PREC_FP32, PREC_BF16, PREC_TF32 = 0, 1, 2
VARIANT_MULTI_INPUT, VARIANT_UNIT_ONLY = 0, 1
F32, F16, BF16 = 0, 1, 2


class Config(C.Structure):
    _fields_ = [
        ("variant", C.c_int32), ("precision", C.c_int32), ("n_ups", C.c_int32),
        ("up_rates", C.c_int32 * L2S_MAX_UPS), ("up_ksizes", C.c_int32 * L2S_MAX_UPS),
        ("up_init_ch", C.c_int32), ("n_rk", C.c_int32), ("rk_sizes", C.c_int32 * L2S_MAX_RK),
        ("n_dil", C.c_int32), ("rk_dils", (C.c_int32 * L2S_MAX_DIL) * L2S_MAX_RK),
        ("num_embeddings", C.c_int32), ("embedding_dim", C.c_int32), ("num_mels", C.c_int32),
        ("spk_dim", C.c_int32), ("num_speakers", C.c_int32), ("multispkr", C.c_int32),
        ("model_in_dim", C.c_int32),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("inp", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("out_raw", C.c_void_p),
        ("out_act", C.c_void_p), ("res", C.c_void_p), ("acc_in", C.c_void_p),
        ("act_bf16", C.c_int32), ("batch", C.c_int32), ("lin", C.c_int32), ("cin_pad", C.c_int32),
        ("ntaps", C.c_int32), ("ntot", C.c_int32), ("mrows", C.c_int32),
        ("tap_off", C.c_int32 * 16), ("out_shift", C.c_int64), ("out_valid", C.c_int64),
        ("scale", C.c_float), ("slope", C.c_float),
    ]


EXPORTS = [
    "l2s_create", "l2s_destroy", "l2s_set_weight", "l2s_finalize", "l2s_workspace_bytes", "l2s_hop",
    "l2s_forward", "l2s_forward_i16", "l2s_poll_index_error", "l2s_launch_count", "l2s_last_error",
    "l2s_version", "l2s_debug_tap", "l2s_debug_conv", "l2s_debug_set", "l2s_debug_layer_time",
    "l2s_io_read_npy_f32", "l2s_io_write_wav_i16", "l2s_io_units_to_ids",            # include/l2s_hand_off.h
]
IO_OK, IO_ERR_ARG, IO_ERR_OPEN, IO_ERR_FORMAT = range(4)
IO_REQUIRE_1D, IO_REQUIRE_2D, IO_REQUIRE_F32 = 1, 2, 4

_lib = None


def load():
    """dlopen the in-tree library and declare every prototype of l2s_vocoder.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python lip2speech-unit_b200/build.py` "
            "(there is no CPU / PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.l2s_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.l2s_create.restype = C.c_int
    lib.l2s_destroy.argtypes = [vp]
    lib.l2s_destroy.restype = None
    lib.l2s_set_weight.argtypes = [vp, C.c_char_p, vp, i64]
    lib.l2s_set_weight.restype = C.c_int
    lib.l2s_finalize.argtypes = [vp, C.c_int]
    lib.l2s_finalize.restype = C.c_int
    lib.l2s_workspace_bytes.argtypes = [vp, i32, i32]
    lib.l2s_workspace_bytes.restype = i64
    lib.l2s_hop.argtypes = [vp]
    lib.l2s_hop.restype = i32
    lib.l2s_forward.argtypes = [vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, vp, i64]
    lib.l2s_forward.restype = C.c_int
    lib.l2s_forward_i16.argtypes = [vp, vp, vp, vp, i32, vp, i32, i32, i32, vp, vp, vp, i64]
    lib.l2s_forward_i16.restype = C.c_int
    lib.l2s_poll_index_error.argtypes = [vp]
    lib.l2s_poll_index_error.restype = C.c_int
    lib.l2s_launch_count.argtypes = [vp, i32, i32]
    lib.l2s_launch_count.restype = i32
    lib.l2s_last_error.argtypes = [vp]
    lib.l2s_last_error.restype = C.c_char_p
    lib.l2s_version.argtypes = []
    lib.l2s_version.restype = C.c_char_p
    lib.l2s_debug_tap.argtypes = [vp, C.c_char_p, vp, i64]
    lib.l2s_debug_tap.restype = C.c_int
    lib.l2s_debug_conv.argtypes = [C.POINTER(ConvDesc), i32, i32, vp, C.c_char_p, i32]
    lib.l2s_debug_conv.restype = C.c_int
    lib.l2s_debug_layer_time.argtypes = [vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_char_p, i32]
    lib.l2s_debug_layer_time.restype = C.c_int
    lib.l2s_debug_set.argtypes = [C.c_char_p, i64]
    lib.l2s_debug_set.restype = C.c_int
    pp, ip = C.POINTER(C.c_char_p), C.POINTER(i32)
    lib.l2s_io_read_npy_f32.argtypes = [pp, i32, vp, i64, ip, i32, i32, ip, i32, ip]
    lib.l2s_io_read_npy_f32.restype = C.c_int
    lib.l2s_io_write_wav_i16.argtypes = [pp, i32, vp, i64, ip, i32, i32, ip]
    lib.l2s_io_write_wav_i16.restype = C.c_int
    lib.l2s_io_units_to_ids.argtypes = [pp, i32, pp, i32, vp, i64, ip, ip, i32]
    lib.l2s_io_units_to_ids.restype = C.c_int
    _lib = lib
    return lib


def last_error(lib, handle) -> str:
    s = lib.l2s_last_error(handle)
    return s.decode("utf-8", "replace") if s else ""


def raise_for(lib, handle, status: int):
    """Map a C status onto the exception type the reference raises for the same
    condition (SURVEY.md 8b 'Error conventions')."""
    if status == OK:
        return
    msg = last_error(lib, handle)
    if status == ERR_SHAPE:
        raise RuntimeError(msg)               # torch.cat size mismatch, models_multi_input.py:73
    if status == ERR_INDEX:
        raise IndexError(msg)                 # embedding index out of range
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(f"l2s_vocoder status {status}: {msg}")
