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


class _SctLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predictions, label, c, eps):
        require_cuda(predictions, "predictions")
        B = int(label.numel())
        K = int(predictions.numel()) // max(B, 1)
        Yv = predictions.detach().reshape(B, K).float().contiguous()
        lab = label.detach().reshape(B).to(device=Yv.device, dtype=torch.int64).contiguous()
        cen = c.detach().reshape(B).to(device=Yv.device, dtype=torch.float32).contiguous()
        loss = torch.empty(B, dtype=torch.float32, device=Yv.device)
        dY = torch.empty_like(Yv)
        _lib.call("mpo_sct_loss", _ptr(Yv), _ptr(lab), _ptr(cen), ctypes.c_float(eps), ctypes.c_float(1.0), _ptr(loss),
                  _ptr(dY), B, K, _stream())
        ctx.save_for_backward(dY)
        ctx.shape = predictions.shape
        return loss.reshape(())        # the reference returns a 0-d tensor for its single slide (loss.py:75-85)

    @staticmethod
    def backward(ctx, g):
        (dY,) = ctx.saved_tensors
        return (dY * g).reshape(ctx.shape), None, None, None


class SurvivalClassificationTobitLoss:
    """reference: models/loss.py:62-85.  predictions: the soft-maxed class probabilities Y [1, 4] of one slide."""

    def __call__(self, predictions, label, c, eps: float = 1e-7):
        if int(label.numel()) != 1:
            raise RuntimeError("SurvivalClassificationTobitLoss takes one slide per call, as in the reference")
        return _SctLossFn.apply(predictions, label, c, float(eps))


class _AttnNormFn(torch.autograd.Function):
    """lambda_reg * ||attention||_2 over the whole map of one slide (torch.norm(attention, p=2), loss.py:97)."""

    @staticmethod
    def forward(ctx, attention, lambda_reg):
        require_cuda(attention, "attention")
        from . import bagpass as bp
        # the norm does not depend on the layout: any tensor (the reference's own test feeds a [6,10,10] one) is taken
        # flat and folded into the library's [6][columns] map layout; a contiguous [6, N] map is used where it lies
        flat = attention.detach().to(torch.float32).reshape(-1)
        pad = (-flat.numel()) % bp.Q
        if pad:
            flat = torch.cat([flat, flat.new_zeros(pad)])
        A = flat.reshape(bp.Q, -1).contiguous()
        rows = A.shape[1]
        info, prefix = bp._tile_table_np((rows,))
        dev = A.device
        tile_info = torch.from_numpy(info).to(dev)
        tile_prefix = torch.from_numpy(prefix).to(dev)
        bag = _lib.MpoBag(0, rows, tile_info.data_ptr(), tile_prefix.data_ptr(), int(info.shape[0]), 1)
        sumsq = torch.empty((1, bp.Q), dtype=torch.float32, device=dev)
        _lib.call("mpo_attn_map_dot", ctypes.byref(bag), _ptr(A), _ptr(A), _ptr(sumsq), _stream())
        reg = torch.empty(1, dtype=torch.float32, device=dev)
        dA = torch.empty_like(A)
        _lib.call("mpo_cesar_reg", ctypes.byref(bag), _ptr(A), _ptr(sumsq), ctypes.c_float(lambda_reg), ctypes.c_float(1.0),
                  _ptr(reg), _ptr(dA), _stream())
        ctx.save_for_backward(dA)
        ctx.shape, ctx.numel = attention.shape, attention.numel()
        return reg.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dA,) = ctx.saved_tensors
        return (dA.reshape(-1)[:ctx.numel] * g).reshape(ctx.shape), None


class CrossEntropySurvivalAttnRegLoss:
    """reference: models/loss.py:88-101 -- CES + lambda_reg * ||attention||_2; returns (loss, attn_loss).  The gradient
    of the norm flows into the co-attention map and from there through the bag backward pass (mpo_bag_bwd's d_amap)."""

    def __init__(self, alpha=0.75, eps=1e-7, lambda_reg=0.01):
        self.alpha = alpha
        self.eps = eps
        self.lambda_reg = lambda_reg
        self.ces = CrossEntropySurvivalLoss(self.alpha, self.eps)

    def __call__(self, hazards, S, Y, c, attention):
        loss = self.ces(hazards, S, Y, c)
        attn_loss = _AttnNormFn.apply(attention, float(self.lambda_reg))
        loss = (loss + attn_loss).mean()
        return loss, attn_loss


class CoxSurvivalLoss:
    """reference: models/loss.py:46-59 -- a batch-level partial likelihood built with Python loops over numpy; no
    driver of the reference selects it (models/*/main.py wire ce / ces / sct / cesar) and it has no device work."""

    def __call__(self, hazards, S, c):
        raise NotImplementedError("CoxSurvivalLoss is not part of the B200 slide path: none of the reference's drivers "
                                  "select it; use ces / nll / sct / cesar")
