"""ctypes binding of libltb200 (include/ltb200.h).  Fails loudly when the CUDA library is missing —
there is deliberately no CPU fallback behind this module."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libltb200.so")

LTB_SESSION_KEEP_LAYERS = 1
LTB_SESSION_NO_GRAPH = 2
LTB_SESSION_NO_HALO = 4


class LtbError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("N", "IH", "IW", "Cin", "Cout", "KH", "KW", "sy", "sx", "pad", "transposed", "relu", "has_res",
                 "force_path")]


_SIGS = {
    "ltb_version": (C.c_int, []),
    "ltb_last_error": (C.c_char_p, []),
    "ltb_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ltb_set_device": (C.c_int, [C.c_int]),
    "ltb_w2l_model_create": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ltb_w2l_model_create_from_device": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ltb_w2l_model_destroy": (C.c_int, [C.c_void_p]),
    "ltb_w2l_avatar_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_void_p)]),
    "ltb_w2l_avatar_destroy": (C.c_int, [C.c_void_p]),
    "ltb_w2l_session_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_void_p)]),
    "ltb_w2l_session_destroy": (C.c_int, [C.c_void_p]),
    "ltb_w2l_mel_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ltb_w2l_infer": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ltb_w2l_paste": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ltb_w2l_paste_pred": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ltb_w2l_paste_batch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ltb_w2l_mel_resident": (C.c_int, [C.c_void_p]),
    "ltb_w2l_step_async": (C.c_int, [C.c_void_p, C.c_int]),
    "ltb_w2l_profile_ops": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]),
    "ltb_w2l_sync": (C.c_int, [C.c_void_p]),
    "ltb_w2l_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ltb_w2l_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong)]),
    "ltb_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "ltb_host_free": (C.c_int, [C.c_void_p]),
    "ltb_w2l_num_layers": (C.c_int, []),
    "ltb_w2l_layer_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ltb_w2l_layer_read": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "ltb_umma_probe": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ltb_conv2d_f16": (C.c_int, [C.POINTER(ConvDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def lib() -> C.CDLL:
    """Load libltb200.so (once).  Raises LtbError if it has not been built — no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LtbError(f"{LIB_PATH} not found: build it with `python -m livetalking_b200.build` "
                           "(sm_100a CUDA library; there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().ltb_last_error()
        raise LtbError(msg.decode("utf-8", "replace") if msg else f"libltb200 error {rc}")
