"""Stand-alone CUDA operators behind the block / fusion classes called on their own (mpo_op_* of include/mpo_b200.h).

The reference's building blocks are nn.Modules with their own `forward` (models/blocks.py, models/fusion.py), and its
unit tests call them directly.  Inside MCAT / NaCAGaT they run fused in the slide pass; called on their own they go
through these wrappers: every arithmetic step is a kernel of libmpo_b200.so (torch only allocates the outputs).
Inference only -- the results carry no autograd graph; training goes through the slide modules.  CPU tensors are
refused: there is no fallback."""
import ctypes
import itertools

import torch

from . import _lib
from .bagpass import _ptr, _stream, require_cuda

ACT = {None: 0, "none": 0, "relu": 1, "elu": 2, "tanh": 3, "sigmoid": 4}
OP_ADD, OP_MUL, OP_ACT = 0, 1, 16
_site = itertools.count(1000)          # dropout sites of stand-alone calls: outside the slide path's site numbers


def _seed():
    from .slidepath import _next_seed
    return _next_seed()


def _prep(x, what):
    require_cuda(x, what)
    return x.detach().to(torch.float32).contiguous()


def linear(x, weight, bias=None, act=None, drop_p=0.0):
    """dropout(act(x W^T + b)) for x [rows, in] (or [in])."""
    x = _prep(x, "input")
    w = _prep(weight, "weight")
    squeeze = x.dim() == 1
    x2 = x.reshape(1, -1) if squeeze else x.reshape(-1, x.shape[-1])
    if x2.shape[1] != w.shape[1]:
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx%d and %dx%d)"
                           % (x2.shape[0], x2.shape[1], w.shape[1], w.shape[0]))
    b = _prep(bias, "bias") if bias is not None else None
    y = torch.empty((x2.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
    _lib.call("mpo_op_linear", _ptr(x2), x2.shape[1], _ptr(w), _ptr(b), _ptr(y), y.shape[1], x2.shape[0], w.shape[1],
              w.shape[0], ACT[act], ctypes.c_float(float(drop_p)), _seed() if drop_p > 0 else 0, next(_site) & 0xFFFF,
              _stream())
    return y.reshape(-1) if squeeze else y.reshape(x.shape[:-1] + (w.shape[0],))


def linear_into(x2, weight, bias, y, ldy, act=None, drop_p=0.0):
    """the same, rows of x2 [rows, in] written into a wider row-major buffer y with row pitch ldy."""
    w, b = _prep(weight, "weight"), _prep(bias, "bias")
    _lib.call("mpo_op_linear", _ptr(x2), x2.shape[1], _ptr(w), _ptr(b), _ptr(y), ldy, x2.shape[0], w.shape[1], w.shape[0],
              ACT[act], ctypes.c_float(float(drop_p)), _seed() if drop_p > 0 else 0, next(_site) & 0xFFFF, _stream())


def layernorm(x, weight, bias, eps=1e-5):
    x = _prep(x, "input")
    y = torch.empty_like(x)
    cols = x.shape[-1]
    _lib.call("mpo_op_layernorm", _ptr(x), _ptr(_prep(weight, "weight")), _ptr(_prep(bias, "bias")), _ptr(y),
              x.numel() // cols, cols, ctypes.c_float(eps), _stream())
    return y


def _ewise(op, a, b=None):
    a = _prep(a, "input")
    if b is not None:
        b = _prep(b, "input")
        if b.shape != a.shape:
            raise RuntimeError("element-wise operands differ in shape: %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    y = torch.empty_like(a)
    _lib.call("mpo_op_ewise", op, _ptr(a), _ptr(b), _ptr(y), a.numel(), _stream())
    return y


def add(a, b):
    return _ewise(OP_ADD, a, b)


def mul(a, b):
    return _ewise(OP_MUL, a, b)


def act(x, kind):
    return _ewise(OP_ACT + ACT[kind], x)


def rowscale(x, g):
    """x [rows, cols] (or [cols]) times one scalar per row (g [rows] / [rows, 1] / [1])."""
    x = _prep(x, "input")
    g = _prep(g, "gate").reshape(-1)
    x2 = x.reshape(1, -1) if x.dim() == 1 else x.reshape(-1, x.shape[-1])
    if g.numel() != x2.shape[0]:
        raise RuntimeError("one gate per row expected")
    y = torch.empty_like(x2)
    _lib.call("mpo_op_rowscale", _ptr(x2), _ptr(g), _ptr(y), x2.shape[0], x2.shape[1], _stream())
    return y.reshape(x.shape)


def dropout(x, p):
    x = _prep(x, "input")
    if p <= 0:
        return x
    y = torch.empty_like(x)
    _lib.call("mpo_op_dropout", _ptr(x), _ptr(y), x.numel(), ctypes.c_float(float(p)), _seed(), next(_site) & 0xFFFF, _stream())
    return y
