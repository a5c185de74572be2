"""
ctypes binding of libbild_b200.so (the C ABI in include/bild_b200.h).

There is no CPU fallback: if the library is missing or no CUDA device is present, every compute entry
point raises.  (The reference falls back to its pure-Python twin with a warning,
/root/reference/bild/cython_imports.py:3-7; this engine deliberately does not.)
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbild_b200.so")

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_uint8_p = ctypes.POINTER(ctypes.c_uint8)
c_uint32_p = ctypes.POINTER(ctypes.c_uint32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class AmisReq(ctypes.Structure):
    """`bildk_amis_req` of include/bild_b200.h: one trajectory's fused AMIS step."""
    _fields_ = [("ens", ctypes.c_void_p), ("ss", c_double_p), ("thetas", c_int64_p), ("A_cur", c_double_p),
                ("logp_cur", c_double_p), ("head", c_double_p), ("per_sample", c_double_p)]


# every symbol include/bild_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "bildk_version": (ctypes.c_int, []),
    "bildk_last_error": (ctypes.c_char_p, []),
    "bildk_device_count": (ctypes.c_int, []),
    "bildk_model_create": (ctypes.c_int, [ctypes.c_int] * 3 + [c_double_p] * 6 + [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bildk_model_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bildk_traj_create": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p, c_uint32_p,
                                          ctypes.POINTER(ctypes.c_void_p)]),
    "bildk_traj_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bildk_logl_runs": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_int32_p, c_uint8_p, c_double_p]),
    "bildk_logl_st": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.POINTER(ctypes.c_int64), c_double_p]),
    "bildk_logl_states": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_int32_p, c_double_p]),
    "bildk_logl_runs_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "bildk_logl_runs_multi": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), c_int32_p, ctypes.c_int, c_int32_p,
                                              c_uint8_p, c_double_p]),
    "bildk_logl_runs_multi_submit": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p), c_int32_p, ctypes.c_int, c_int32_p,
                                                     c_uint8_p, c_double_p, ctypes.POINTER(AmisReq), ctypes.POINTER(ctypes.c_void_p)]),
    "bildk_logl_wait": (ctypes.c_int, [ctypes.c_void_p]),
    "bildk_logl_ready": (ctypes.c_int, [ctypes.c_void_p]),
    "bildk_amis_weights": (ctypes.c_int, [ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_double, c_double_p,
                                           c_double_p, ctypes.c_int]),
    "bildk_marginal_posterior": (ctypes.c_int, [ctypes.c_int] * 4 + [c_int32_p, c_uint8_p, c_double_p, c_double_p, ctypes.c_int]),
    "bildk_amis_log_proposal": (ctypes.c_int, [ctypes.c_int] * 4 + [c_double_p, c_double_p, c_uint8_p, c_double_p,
                                                ctypes.POINTER(ctypes.c_int64), c_double_p]),
    "bildk_choice_pick": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.c_double, c_int64_p]),
    "bildk_choice_dn": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_double, c_int64_p]),
    "bildk_amis_weights_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "bildk_amis_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_uint8_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bildk_amis_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bildk_amis_size": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "bildk_amis_step": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.POINTER(ctypes.c_int64), c_double_p, c_double_p,
                                        c_double_p, c_double_p, c_double_p]),
    "bildk_launch_count": (ctypes.c_longlong, []),
    "bildk_describe_plan": (ctypes.c_char_p, [ctypes.c_void_p, ctypes.c_int]),
    "bildk_debug_tables": (ctypes.c_int, [ctypes.c_int] * 4 + [c_uint8_p]),
}

_lib = None


class BildkError(RuntimeError):
    pass


def load():
    """Load the library (once) and declare all prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BildkError(f"{LIB_PATH} not found - build it with `python -m bild_b200.build` "
                             "(there is no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    """Turn a negative return code into the matching Python exception."""
    if rc == 0:
        return
    msg = load().bildk_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise MemoryError(msg)
    raise BildkError(f"bild_b200 error {rc}: {msg}")


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a, typ):
    return a.ctypes.data_as(typ)
