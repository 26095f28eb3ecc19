"""ctypes binding of libmpo_b200.so (the C ABI declared in include/mpo_b200.h).

The library is the product: if it is missing, or a call fails, the caller gets a RuntimeError --
there is no PyTorch or CPU fallback behind any of these functions.
"""
import ctypes
import os
import threading

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


class MpoLin(ctypes.Structure):
    _fields_ = [("w", c_void_p), ("b", c_void_p), ("gw", c_void_p), ("gb", c_void_p)]


class MpoNorm(ctypes.Structure):
    _fields_ = [("g", c_void_p), ("b", c_void_p), ("gg", c_void_p), ("gb", c_void_p)]


class MpoEncoderLayer(ctypes.Structure):
    _fields_ = [("in_proj", MpoLin), ("out_proj", MpoLin), ("linear1", MpoLin), ("linear2", MpoLin),
                ("norm1", MpoNorm), ("norm2", MpoNorm)]


class MpoPoolHead(ctypes.Structure):
    _fields_ = [("att_a", MpoLin), ("att_b", MpoLin), ("att_c", MpoLin), ("rho", MpoLin)]


class MpoCag(ctypes.Structure):
    _fields_ = [("fc1", MpoLin), ("fc2", MpoLin), ("fc3", MpoLin), ("fc_c", MpoLin), ("G", MpoNorm), ("E", MpoNorm)]


class MpoBilinear(ctypes.Structure):
    _fields_ = [("h1", MpoLin), ("z1", MpoLin), ("o1", MpoLin), ("h2", MpoLin), ("z2", MpoLin), ("o2", MpoLin),
                ("fc1", MpoLin), ("fc2", MpoLin)]


class MpoModel(ctypes.Structure):
    """struct mpo_model (include/mpo_b200.h)."""
    _fields_ = [
        ("variant", c_i32), ("fusion", c_i32), ("n_classes", c_i32), ("omic_dims", c_i32 * 6),
        ("H", MpoLin), ("snn", (MpoLin * 2) * 6), ("coattn_in", MpoLin), ("coattn_out", MpoLin), ("cag", MpoCag),
        ("path_tr", MpoEncoderLayer * 2), ("omic_tr", MpoEncoderLayer * 2),
        ("path_pool", MpoPoolHead), ("omic_pool", MpoPoolHead),
        ("fusion0", MpoLin), ("fusion2", MpoLin), ("gate", MpoLin * 2), ("bil", MpoBilinear), ("classifier", MpoLin),
    ]


class MpoTailIo(ctypes.Structure):
    """struct mpo_tail_io (include/mpo_b200.h)."""
    _fields_ = [
        ("num_slides", c_i32), ("omics", c_void_p * 6), ("ws", c_void_p),
        ("qp", c_void_p), ("qk", c_void_p), ("kc", c_void_p), ("pooled", c_void_p), ("suma", c_void_p), ("dsuma", c_void_p),
        ("dpooled", c_void_p),
        ("dqk", c_void_p), ("dkc", c_void_p), ("dtq", c_void_p),
        ("hazards", c_void_p), ("S", c_void_p), ("Y", c_void_p), ("att_path", c_void_p), ("att_omic", c_void_p),
        ("drop_p", c_float), ("seed", c_u32), ("seed_dev", c_void_p), ("train", c_i32),
    ]


class MpoGeModel(ctypes.Structure):
    """struct mpo_ge_model (include/mpo_b200.h)."""
    _fields_ = [("n_classes", c_i32), ("H", MpoLin), ("sa_in", MpoLin), ("sa_out", MpoLin), ("tr", MpoEncoderLayer * 2),
                ("pool", MpoPoolHead), ("classifier", MpoLin)]


class MpoNacagatBwd(ctypes.Structure):
    """struct mpo_nacagat_bwd (include/mpo_b200.h)."""
    _fields_ = [(n, c_void_p) for n in (
        "h_saved", "t_saved", "scores", "pgate", "lse", "pooled", "suma", "pooled_lo", "dpooled", "dsuma", "d_amap", "amap_dot",
        "qk", "qp", "w_k_f16",
        "dz_ws", "dkg_ws", "dg_ws", "part_dqk", "part_dtq", "part_db", "part_dbk", "part_dkc", "dg_max",
        "dqk", "dkc", "dtq", "grad_w_h", "grad_b_h", "grad_w_k", "grad_b_k")] + [
        ("drop_p", c_float), ("attn_drop_p", c_float), ("seed", c_u32), ("seed_dev", c_void_p)]


class MpoPeerGroup(ctypes.Structure):
    """struct mpo_peer_group (include/mpo_b200.h)."""
    _fields_ = [("world", c_i32), ("rank", c_i32), ("data", c_void_p * 8), ("flags", c_void_p * 8),
                ("grad", c_void_p * 8), ("param", c_void_p * 8), ("epochs", c_void_p)]


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
                    c_void_p, c_void_p, c_void_p, c_u32, c_void_p, c_float, c_void_p],
    "mpo_cast_f16": [c_void_p, c_void_p, c_i64, c_void_p],
    "mpo_bag_gate_fwd": [ctypes.POINTER(MpoBag)] + [c_void_p] * 17 + [c_u32, c_void_p, c_float, c_void_p],
    "mpo_attn_map_dropout": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_u32, c_void_p, c_float, c_void_p],
    "mpo_advance_seed": [c_void_p, c_void_p],
    "mpo_attn_map": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_void_p],
    "mpo_bag_bwd": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p],
    "mpo_attn_map_dot": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_void_p],
    "mpo_cesar_reg": [ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_void_p],
    "mpo_sct_loss": [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_i32, c_i32, c_void_p],
    "mpo_l1_sum": [c_void_p, c_i64, c_void_p, c_void_p],
    "mpo_l1_grad": [c_void_p, c_void_p, c_i64, c_float, c_void_p],
    "mpo_peer_alloc": [c_i64, ctypes.POINTER(c_void_p)],
    "mpo_peer_free": [c_void_p],
    "mpo_peer_export": [c_void_p, c_void_p],
    "mpo_peer_open": [c_void_p, ctypes.POINTER(c_void_p)],
    "mpo_peer_close": [c_void_p],
    "mpo_peer_warmup": [],
    "mpo_peer_barrier": [ctypes.POINTER(MpoPeerGroup), c_i32, c_void_p],
    "mpo_peer_lse_combine": [ctypes.POINTER(MpoPeerGroup), c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "mpo_peer_adam_step": [ctypes.POINTER(MpoPeerGroup), c_i32, c_i64, c_i64, c_void_p, c_void_p, c_i64, c_float, c_float,
                           c_float, c_float, c_float, c_float, c_void_p, c_i32, c_void_p],
    "mpo_bag_bwd_nacagat": [ctypes.POINTER(MpoBag), ctypes.POINTER(MpoNacagatBwd), c_void_p],
    "mpo_adam_step": [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_float, c_float, c_float, c_float, c_float,
                      c_void_p, c_i32, c_void_p],
    "mpo_tail_side_adam": [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_float, c_float, c_float, c_float, c_float,
                           c_void_p, c_i32, c_void_p],
    "mpo_ge_fwd": [ctypes.POINTER(MpoGeModel), c_i64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                   c_u32, c_i32, c_void_p],
    "mpo_ge_ce_loss": [c_void_p, c_void_p, c_i32, c_float, c_void_p, c_void_p, c_void_p],
    "mpo_ge_bwd": [ctypes.POINTER(MpoGeModel), ctypes.POINTER(MpoBag), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                   c_void_p, c_void_p, c_float, c_float, c_u32, c_i32, c_void_p],
    "mpo_lse_combine": [c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_void_p],
    "mpo_tail_pre_fwd": [ctypes.POINTER(MpoModel), ctypes.POINTER(MpoTailIo), c_void_p],
    "mpo_tail_post_fwd": [ctypes.POINTER(MpoModel), ctypes.POINTER(MpoTailIo), c_void_p],
    "mpo_surv_loss": [c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_void_p,
                      c_void_p, c_i32, c_i32, c_void_p],
    "mpo_tail_post_bwd": [ctypes.POINTER(MpoModel), ctypes.POINTER(MpoTailIo), c_void_p, c_void_p, c_void_p, c_void_p],
    "mpo_tail_pre_bwd": [ctypes.POINTER(MpoModel), ctypes.POINTER(MpoTailIo), c_void_p],
    "mpo_tail_post_step": [ctypes.POINTER(MpoModel), ctypes.POINTER(MpoTailIo), c_i32, c_void_p, c_void_p, c_float, c_float,
                           c_float, c_void_p, c_void_p, c_void_p, c_i32, c_void_p],
    # stand-alone operators (the block / fusion classes called on their own)
    "mpo_op_linear": [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_float, c_u32, c_u32,
                      c_void_p],
    "mpo_op_layernorm": [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_float, c_void_p],
    "mpo_op_ewise": [c_i32, c_void_p, c_void_p, c_void_p, c_i64, c_void_p],
    "mpo_op_rowscale": [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_void_p],
    "mpo_op_dropout": [c_void_p, c_void_p, c_i64, c_float, c_u32, c_u32, c_void_p],
    "mpo_op_bil_gate": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_void_p],
    "mpo_op_bil_kron": [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_float, c_u32, c_u32, c_void_p],
}
# functions with a non-int return type
OTHER_EXPORTS = ["mpo_tail_ws_floats", "mpo_tail_ws_lookup", "mpo_sizeof", "mpo_launch_count", "mpo_ge_ws_floats",
                 "mpo_peer_exchange_bytes", "mpo_peer_flags_offset", "mpo_peer_slice"]


def _declare(L):
    for name, argtypes in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = c_int
        fn.argtypes = argtypes
    L.mpo_tail_ws_floats.restype = c_i64
    L.mpo_tail_ws_floats.argtypes = [ctypes.POINTER(MpoModel), c_i32]
    L.mpo_ge_ws_floats.restype = c_i64
    L.mpo_ge_ws_floats.argtypes = [c_i64]
    L.mpo_launch_count.restype = c_i64
    L.mpo_launch_count.argtypes = [c_i32]
    L.mpo_sizeof.restype = c_i64
    L.mpo_sizeof.argtypes = [c_i32]
    L.mpo_peer_exchange_bytes.restype = c_i64
    L.mpo_peer_exchange_bytes.argtypes = []
    L.mpo_peer_flags_offset.restype = c_i64
    L.mpo_peer_flags_offset.argtypes = []
    L.mpo_peer_slice.restype = None
    L.mpo_peer_slice.argtypes = [c_i64, c_i64, c_i32, c_i32, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]
    for which, st in enumerate((MpoBag, MpoModel, MpoTailIo, MpoNacagatBwd, MpoGeModel, MpoPeerGroup)):
        if L.mpo_sizeof(which) != ctypes.sizeof(st):
            raise RuntimeError("ABI mismatch: %s is %d bytes in libmpo_b200.so but %d in the ctypes binding"
                               % (st.__name__, L.mpo_sizeof(which), ctypes.sizeof(st)))
    L.mpo_tail_ws_lookup.restype = c_i64
    L.mpo_tail_ws_lookup.argtypes = [ctypes.POINTER(MpoModel), c_i32, ctypes.c_char_p, ctypes.POINTER(c_i64)]


def exported_symbols():
    """Names include/mpo_b200.h declares (used by the CPU-side ABI test)."""
    return ["mpo_last_error", "mpo_version"] + list(SIGNATURES.keys()) + OTHER_EXPORTS


_CALL_LOCK = threading.Lock()


def call(name, *args):
    """One C-ABI call; non-zero return codes become RuntimeError.  The library keeps per-process launch state (parameter
    blocks, tensor-map cache, last error), and ctypes drops the GIL for the duration of a call, so calls from several
    Python threads (e.g. nn.DataParallel's replica threads) are serialised here; they only enqueue work on streams."""
    L = lib()
    with _CALL_LOCK:
        rc = getattr(L, name)(*args)
        if rc != 0:
            msg = L.mpo_last_error().decode("utf-8", "replace")
            raise RuntimeError("%s failed (%d): %s" % (name, rc, msg))
