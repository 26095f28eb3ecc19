"""Discrete-time survival losses on the device (reference: models/loss.py).

Same call signatures as the reference classes; the arithmetic (loss value and its gradient w.r.t. hazards and S)
is the surv_loss kernel of the slide tail, reached through mpo_surv_loss.  Inputs must be CUDA tensors."""
import ctypes

import torch

from . import _lib
from .bagpass import _ptr, _stream, require_cuda

NLL, CES = 0, 1


class _SurvLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hazards, S, Y, c, kind, alpha, eps):
        require_cuda(hazards, "hazards")
        B = int(Y.numel())
        K = int(hazards.shape[-1])
        hz = hazards.detach().reshape(B, K).float().contiguous()
        Sv = S.detach().reshape(B, K).float().contiguous()
        # labels that are still on the host are range-checked here for free; for device labels a check would cost
        # two device->host syncs per slide (the host could no longer run ahead of the GPU), so the kernel guards
        # instead: an out-of-range label reads nothing and poisons that slide's loss with NaN
        if not Y.is_cuda and B > 0 and (int(Y.min()) < 0 or int(Y.max()) >= K):
            raise IndexError("survival label out of range [0, %d)" % K)
        lab = Y.detach().reshape(B).to(device=hz.device, dtype=torch.int64).contiguous()
        cen = c.detach().reshape(B).to(device=hz.device, dtype=torch.float32).contiguous()
        loss = torch.empty(B, dtype=torch.float32, device=hz.device)
        dhz = torch.empty_like(hz)
        dS = torch.empty_like(Sv)
        _lib.call("mpo_surv_loss", kind, _ptr(hz), _ptr(Sv), _ptr(lab), _ptr(cen), ctypes.c_float(alpha),
                  ctypes.c_float(eps), ctypes.c_float(1.0 / B), _ptr(loss), _ptr(dhz), _ptr(dS), B, K, _stream())
        ctx.save_for_backward(dhz, dS)
        ctx.shapes = (hazards.shape, S.shape)
        return loss.mean()

    @staticmethod
    def backward(ctx, g):
        dhz, dS = ctx.saved_tensors
        return (dhz * g).reshape(ctx.shapes[0]), (dS * g).reshape(ctx.shapes[1]), None, None, None, None, None


class NegativeLogLikelihoodSurvivalLoss:
    """reference: models/loss.py:31-43."""

    def __call__(self, hazards, S, Y, c, alpha=0.15, eps=1e-7):
        return _SurvLossFn.apply(hazards, S, Y, c, NLL, float(alpha), float(eps))


class CrossEntropySurvivalLoss:
    """reference: models/loss.py:5-28."""

    def __init__(self, alpha=0.75, eps=1e-7):
        self.alpha = alpha
        self.eps = eps

    def __call__(self, hazards, S, Y, c):
        return _SurvLossFn.apply(hazards, S, Y, c, CES, float(self.alpha), float(self.eps))
