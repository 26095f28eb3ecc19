"""ctypes binding of libmpo_b200.so (the C ABI declared in include/mpo_b200.h).

The library is the product: if it is missing, or a call fails, the caller gets a RuntimeError --
there is no PyTorch or CPU fallback behind any of these functions.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmpo_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_u32 = ctypes.c_uint32
c_float = ctypes.c_float


class MpoBag(ctypes.Structure):
    """struct mpo_bag (include/mpo_b200.h)."""
    _fields_ = [
        ("x", c_void_p),
        ("total_rows", c_i64),
        ("tile_info", c_void_p),
        ("tile_prefix", c_void_p),
        ("num_tiles", c_i32),
        ("num_slides", c_i32),
    ]


_lib = None


def build_hint():
    return ("libmpo_b200.so not found at %s -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or multimodal-path-omic_b200/csrc/build.sh). "
            "There is no CPU/PyTorch fallback for this path." % LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(build_hint())
        L = ctypes.CDLL(LIB_PATH)
        L.mpo_last_error.restype = ctypes.c_char_p
        L.mpo_last_error.argtypes = []
        L.mpo_version.restype = c_int
        _declare(L)
        _lib = L
    return _lib


# name -> argtypes; every function returns int (0 = ok)
SIGNATURES = {
    "mpo_cast_bf16": [c_void_p, c_void_p, c_i64, c_void_p],
    "mpo_bag_fwd": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                    c_void_p, c_void_p, c_u32, c_float, c_void_p],
    "mpo_attn_map": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_void_p],
    "mpo_bag_bwd": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p],
    "mpo_lse_combine": [c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_void_p],
}


def _declare(L):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = c_int
        fn.argtypes = argtypes


def exported_symbols():
    """Names include/mpo_b200.h declares (used by the CPU-side ABI test)."""
    return ["mpo_last_error", "mpo_version"] + list(SIGNATURES.keys())


def call(name, *args):
    L = lib()
    rc = getattr(L, name)(*args)
    if rc != 0:
        msg = L.mpo_last_error().decode("utf-8", "replace")
        raise RuntimeError("%s failed (%d): %s" % (name, rc, msg))
