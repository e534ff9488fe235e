"""ctypes binding of libgpkl.so (C ABI declared in include/gpkl.h).  No CPU fallback exists: if the
library is missing or does not load, importing the ops fails loudly."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpkl.so")

KERNELS = {"rbf": 0, "cauchy": 1}
POSTERIORS = {"gp": 0, "diag": 1, "bidiag": 2}
TIERS = {"auto": 0, "generic": 1, "warp": 2, "block": 3}
FLAG_GRAD_ELL_P = 1
FLAG_PER_PAIR_PRIOR = 2
FLAG_PHILOX_EPS = 4

# every symbol include/gpkl.h declares (tests check the library exports exactly these)
SYMBOLS = ("gpkl_version", "gpkl_strerror", "gpkl_workspace_bytes", "gpkl_forward", "gpkl_backward",
           "gpkl_step_host_bytes", "gpkl_step_host", "gpkl_launch_count", "gpkl_profile_enable",
           "gpkl_profile_read", "gpkl_fp32_peak_launch", "gpkl_recon_workspace_bytes", "gpkl_recon_forward",
           "gpkl_recon_backward", "gpkl_recog_workspace_bytes", "gpkl_recog_forward", "gpkl_recog_backward",
           "gpkl_collate_workspace_bytes", "gpkl_collate", "gpkl_impute_workspace_bytes", "gpkl_impute", "gpkl_philox_normal")


class GpklDesc(ctypes.Structure):
    _fields_ = [("B", ctypes.c_int32), ("D", ctypes.c_int32), ("T_max", ctypes.c_int32), ("S", ctypes.c_int32),
                ("total_T", ctypes.c_int64), ("kernel", ctypes.c_int32), ("posterior", ctypes.c_int32),
                ("noise", ctypes.c_float), ("flags", ctypes.c_int32), ("tier", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "gpkl: %s not found -- the CUDA library is the only implementation of this path (no CPU "
            "fallback).  Build it with `python gp-vae_b200/build.py` or `python -c 'import __graft_entry__ as g; "
            "g.build()'`." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    dp = ctypes.POINTER(GpklDesc)
    L.gpkl_version.restype = i32
    L.gpkl_strerror.restype = ctypes.c_char_p
    L.gpkl_strerror.argtypes = [i32]
    L.gpkl_workspace_bytes.restype = sz
    L.gpkl_workspace_bytes.argtypes = [dp]
    L.gpkl_forward.restype = i32
    L.gpkl_forward.argtypes = [dp] + [vp] * 12 + [vp, sz, vp]
    L.gpkl_backward.restype = i32
    L.gpkl_backward.argtypes = [dp] + [vp] * 15 + [vp, sz, vp]
    L.gpkl_step_host_bytes.restype = sz
    L.gpkl_step_host_bytes.argtypes = [dp]
    L.gpkl_step_host.restype = i32
    L.gpkl_step_host.argtypes = [dp] + [vp] * 15 + [vp, sz, vp]
    L.gpkl_launch_count.restype = ctypes.c_int64
    L.gpkl_profile_enable.restype = i32
    L.gpkl_profile_enable.argtypes = [i32]
    L.gpkl_profile_read.restype = i32
    L.gpkl_profile_read.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32),
                                    ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)]
    L.gpkl_fp32_peak_launch.restype = i32
    L.gpkl_fp32_peak_launch.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_double), vp]
    i64 = ctypes.c_int64
    L.gpkl_recon_workspace_bytes.restype = sz
    L.gpkl_recon_workspace_bytes.argtypes = [ctypes.c_int32]
    L.gpkl_recon_forward.restype = i32
    L.gpkl_recon_forward.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, i64, vp, vp, vp, vp, vp, sz, vp]
    L.gpkl_recon_backward.restype = i32
    L.gpkl_recon_backward.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, i64, vp, vp, vp, vp, vp, vp, sz, vp]
    L.gpkl_recog_workspace_bytes.restype = sz
    L.gpkl_recog_workspace_bytes.argtypes = [dp]
    L.gpkl_recog_forward.restype = i32
    L.gpkl_recog_forward.argtypes = [dp] + [vp] * 10 + [vp, sz, vp]
    L.gpkl_recog_backward.restype = i32
    L.gpkl_recog_backward.argtypes = [dp] + [vp] * 13 + [vp, sz, vp]
    L.gpkl_collate_workspace_bytes.restype = sz
    L.gpkl_collate_workspace_bytes.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.gpkl_collate.restype = i32
    L.gpkl_collate.argtypes = [ctypes.c_int32] * 5 + [vp] * 7 + [vp, sz, vp]
    L.gpkl_philox_normal.restype = i32
    L.gpkl_philox_normal.argtypes = [vp, i64, vp, vp]
    L.gpkl_impute_workspace_bytes.restype = sz
    L.gpkl_impute_workspace_bytes.argtypes = [ctypes.c_int32]
    L.gpkl_impute.restype = i32
    L.gpkl_impute.argtypes = [ctypes.c_int32] * 5 + [ctypes.c_float] * 2 + [vp] * 7 + [vp, sz, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().gpkl_strerror(rc).decode())
