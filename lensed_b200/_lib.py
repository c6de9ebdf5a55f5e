"""ctypes binding of the C ABI in include/lensed_cuda.h.

The shared library is built in-tree (``lensed_b200/liblensed_cuda.so``, see
``__graft_entry__.build``).  There is no fallback of any kind: if the library
is missing or a symbol cannot be bound, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblensed_cuda.so")


class LensedCudaError(RuntimeError):
    """A C-ABI call returned a non-zero status; carries ``code`` and the
    library's message (NVRTC log included for build failures)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[lcu error {code}] {message}")
        self.code = code
        self.message = message


class LcuParam(C.Structure):
    _fields_ = [("name", C.c_char * 16), ("type", C.c_int), ("bounds", C.c_float * 2), ("defval", C.c_float)]


class LcuObjectSpec(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ipp", C.POINTER(C.c_int))]


class LcuModelDesc(C.Structure):
    _fields_ = [
        ("width", C.c_size_t), ("height", C.c_size_t), ("pcs", C.c_float * 4),
        ("nq", C.c_size_t), ("qq", C.c_void_p), ("ww", C.c_void_p),
        ("image", C.c_void_p), ("weight", C.c_void_p),
        ("psf", C.c_void_p), ("psf_width", C.c_size_t), ("psf_height", C.c_size_t),
        ("max_batch", C.c_size_t), ("flags", C.c_uint),
    ]


class LcuProfile(C.Structure):
    _fields_ = [("evaluations", C.c_ulonglong), ("upload_ms", C.c_double), ("set_params_ms", C.c_double),
                ("render_ms", C.c_double), ("convolve_ms", C.c_double), ("reduce_ms", C.c_double),
                ("download_ms", C.c_double)]


# every symbol include/lensed_cuda.h declares: (restype, argtypes)
SYMBOLS = {
    "lcu_version": (C.c_int, []),
    "lcu_last_error": (C.c_char_p, []),
    "lcu_launch_count": (C.c_ulonglong, []),
    "lcu_create": (C.c_int, [C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_void_p)]),
    "lcu_destroy": (None, [C.c_void_p]),
    "lcu_object_info": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t), C.POINTER(LcuParam), C.c_size_t]),
    "lcu_object_pairable": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_char_p)]),
    "lcu_quad_rule_count": (C.c_int, []),
    "lcu_quad_rule_name": (C.c_char_p, [C.c_int]),
    "lcu_quad_rule_info": (C.c_char_p, [C.c_int]),
    "lcu_quad_rule": (C.c_int, [C.c_char_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "lcu_model_create": (C.c_int, [C.c_void_p, C.POINTER(LcuObjectSpec), C.c_size_t, C.POINTER(LcuModelDesc),
                                   C.POINTER(C.c_void_p)]),
    "lcu_model_destroy": (None, [C.c_void_p]),
    "lcu_model_npars": (C.c_size_t, [C.c_void_p]),
    "lcu_model_words": (C.c_size_t, [C.c_void_p]),
    "lcu_model_max_batch": (C.c_size_t, [C.c_void_p]),
    "lcu_model_rays_per_thread": (C.c_int, [C.c_void_p]),
    "lcu_model_source": (C.c_char_p, [C.c_void_p]),
    "lcu_model_build_log": (C.c_char_p, [C.c_void_p]),
    "lcu_model_cubin": (C.c_size_t, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "lcu_model_kernel_usage": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "lcu_model_set_rows": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "lcu_model_set_data": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcu_model_make_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_double, C.c_void_p]),
    "lcu_model_get_weight": (C.c_int, [C.c_void_p, C.c_void_p]),
    "lcu_loglike": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "lcu_loglike_async": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "lcu_loglike_wait": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "lcu_loglike_batch": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "lcu_loglike_batch_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcu_render": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcu_set_params": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "lcu_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "lcu_profile_get": (C.c_int, [C.c_void_p, C.POINTER(LcuProfile)]),
    "lcu_measure_fp32_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
}

LCU_FAST_MATH = 1
LCU_OBJ_SHARED = 2
LCU_FAST_INTRINSICS = 4
LCU_NO_PAIR = 8
LCU_FAST_LENS_INTRINSICS = 16
LCU_FAST_ATANH = 32
LCU_SOURCE_ONLY = 64


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lensed_b200/csrc`.  lensed_b200 has no CPU or pure-Python fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int):
    if rc != 0:
        msg = lib.lcu_last_error()
        raise LensedCudaError(rc, msg.decode(errors="replace") if msg else "")
