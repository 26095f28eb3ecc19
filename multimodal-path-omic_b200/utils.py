"""Helpers shared by the parameter containers (reference: models/utils.py)."""
import math

import torch
import torch.nn as nn


def init_max_weights(module):
    """N(0, 1/sqrt(fan_in)) weights and zero biases for every nn.Linear -- reference: models/utils.py:43-48
    (nn.Bilinear keeps torch's default init, exactly as there: the type check is `type(m) == nn.Linear`)."""
    for m in module.modules():
        if type(m) == nn.Linear:
            std = 1.0 / math.sqrt(m.weight.size(1))
            m.weight.data.normal_(0, std)
            m.bias.data.zero_()


class _L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, *params):
        import ctypes
        from . import _lib
        from .bagpass import _ptr, _stream, require_cuda
        dev = params[0].device
        total = torch.zeros(1, dtype=torch.float32, device=dev)
        for p in params:
            require_cuda(p, "l1_reg parameter")
            q = p.detach()
            if q.dtype != torch.float32 or not q.is_contiguous():
                raise RuntimeError("l1_reg expects contiguous float32 parameters")
            _lib.call("mpo_l1_sum", _ptr(q), q.numel(), _ptr(total), _stream())
        ctx.params = params
        return total.reshape(())

    @staticmethod
    def backward(ctx, g):
        # sign(W) * g per parameter (torch.abs backward).  One host read of the scalar upstream gradient (lambda_reg).
        import ctypes
        from . import _lib
        from .bagpass import _ptr, _stream
        scale = float(g)
        grads = []
        for p in ctx.params:
            if not p.requires_grad:
                grads.append(None)
                continue
            out = torch.zeros_like(p)
            _lib.call("mpo_l1_grad", _ptr(p.detach()), _ptr(out), p.numel(), ctypes.c_float(scale), _stream())
            grads.append(out)
        return tuple(grads)


def l1_reg(model):
    """sum over every parameter of sum |W| -- reference: models/utils.py:33-40 (used as `reg_function(model) *
    lambda_reg`, models/mcat/main.py:58-61).  The sums and the sign gradients are the library's kernels."""
    params = [p for p in model.parameters()]
    if not params:
        return None
    return _L1Fn.apply(*params)
