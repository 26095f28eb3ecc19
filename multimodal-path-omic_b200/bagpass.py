"""Host side of the bag stage: packed-bag layout in HBM and thin wrappers over mpo_bag_fwd / mpo_bag_bwd.

Layout: the bf16 patch features of every slide of a batch sit back to back in one [total_rows, 1024] buffer
(no padding rows); a tile table (one int4 per 128-patch tile) tells the kernels which slide a tile belongs to
and how many of its rows are valid, so ragged bags need no per-slide launches.  PyTorch is used only to own the
device memory and the stream.
"""
import ctypes
import functools

import numpy as np
import torch

from . import _lib

D_IN, D, Q, TILE = 1024, 256, 6, 128


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("%s must live on a CUDA device: the B200 slide path has no CPU fallback" % what)


@functools.lru_cache(maxsize=256)
def _tile_table_np(lengths):
    """tile_info [T,4] = (slide, first packed row, valid rows, tile-in-slide); tile_prefix [B+1]."""
    info = []
    prefix = [0]
    row = 0
    for b, n in enumerate(lengths):
        nt = (n + TILE - 1) // TILE
        for j in range(nt):
            info.append((b, row + j * TILE, min(TILE, n - j * TILE), j))
        prefix.append(prefix[-1] + nt)
        row += n
    info_np = np.asarray(info, dtype=np.int32).reshape(-1, 4)
    return info_np, np.asarray(prefix, dtype=np.int32)


def cast_bf16(src, out=None):
    """fp32 -> bf16 through mpo_cast_bf16 (the library's own kernel)."""
    require_cuda(src, "cast_bf16 input")
    src = src.contiguous()
    if src.dtype != torch.float32:
        raise RuntimeError("cast_bf16 expects float32, got %s" % src.dtype)
    if out is None:
        out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    _lib.call("mpo_cast_bf16", _ptr(src), _ptr(out), src.numel(), _stream())
    return out


class PackedBag:
    """A batch of slides as one packed bf16 bag plus its tile table."""

    def __init__(self, x, lengths):
        require_cuda(x, "bag")
        if x.dtype != torch.bfloat16 or x.dim() != 2 or x.shape[1] != D_IN or not x.is_contiguous():
            raise RuntimeError("packed bag must be a contiguous bf16 [rows, 1024] tensor")
        lengths = tuple(int(n) for n in lengths)
        if sum(lengths) != x.shape[0]:
            raise RuntimeError("slide lengths %s do not add up to %d packed rows" % (lengths, x.shape[0]))
        if any(n <= 0 for n in lengths):
            raise RuntimeError("every slide needs at least one patch")
        self.x = x
        self.lengths = lengths
        self.offsets = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
        info, prefix = _tile_table_np(lengths)
        self.num_tiles = int(info.shape[0])
        self.num_slides = len(lengths)
        self.total_rows = int(x.shape[0])
        self.tile_info = torch.from_numpy(info).to(x.device, non_blocking=False)
        self.tile_prefix = torch.from_numpy(prefix).to(x.device, non_blocking=False)
        self._c = _lib.MpoBag(x.data_ptr(), self.total_rows, self.tile_info.data_ptr(), self.tile_prefix.data_ptr(),
                              self.num_tiles, self.num_slides)

    @classmethod
    def from_slides(cls, slides):
        """slides: list of [N_b, 1024] (or [1, N_b, 1024]) tensors, fp32 or bf16, on the GPU."""
        parts, lengths = [], []
        for s in slides:
            require_cuda(s, "slide bag")
            if s.dim() == 3 and s.shape[0] == 1:
                s = s[0]
            if s.dim() != 2 or s.shape[1] != D_IN:
                raise RuntimeError("a slide bag must be [N, 1024], got %s" % (tuple(s.shape),))
            lengths.append(s.shape[0])
            parts.append(s)
        total = sum(lengths)
        if len(parts) == 1 and parts[0].dtype == torch.bfloat16 and parts[0].is_contiguous():
            return cls(parts[0], lengths)          # one bf16 slide: stream it where it lies
        x = torch.empty((total, D_IN), dtype=torch.bfloat16, device=parts[0].device)
        row = 0
        for s, n in zip(parts, lengths):
            dst = x[row:row + n]
            if s.dtype == torch.bfloat16:
                dst.copy_(s)
            elif s.dtype == torch.float32:
                cast_bf16(s, out=dst)
            else:
                raise RuntimeError("slide bags must be float32 or bfloat16, got %s" % s.dtype)
            row += n
        return cls(x, lengths)

    def c(self):
        return ctypes.byref(self._c)

    def slide_rows(self, b):
        return int(self.offsets[b]), int(self.offsets[b + 1])


class BagWorkspace:
    """Per-batch device buffers of the bag stage (caller-owned in the C ABI; allocated here through torch)."""

    def __init__(self, bag, save_h, nacagat=False, save_gate=False):
        dev = bag.x.device
        T, B, R = bag.num_tiles, bag.num_slides, bag.total_rows
        f32 = dict(dtype=torch.float32, device=dev)
        self.scores = torch.empty((Q, R), **f32)
        self.nacagat = nacagat
        self.scores_g = self.pgate = self.t_saved = self.suma = self.h_lo = self.part_pool_lo = self.pooled_lo = None
        if nacagat:
            self.h_lo = torch.empty((R, D), dtype=torch.float16, device=dev)        # gate pass (mpo_bag_gate_fwd): gated scores always; P and tanh(k) only for a backward pass
            self.scores_g = torch.empty((Q, R), **f32)
            self.suma = torch.empty((B, Q), **f32)
            if save_gate:
                self.pgate = torch.empty((Q, R), **f32)
                self.t_saved = torch.empty((R, D), dtype=torch.float16, device=dev)
                self.part_pool_lo = torch.empty((T, Q, D), **f32)     # remainder part of pooled, for the backward's delta
                self.pooled_lo = torch.empty((B, Q, D), **f32)
        self.part_ml = torch.empty((T, 18 if nacagat else 12), **f32)
        self.part_pool = torch.empty((T, Q, D), **f32)
        self.pooled = torch.empty((B, Q, D), **f32)
        self.lse = torch.empty((B, Q), **f32)
        self.h_saved = torch.empty((R, D), dtype=torch.float16, device=dev) if save_h else None
        self.dz = None
        self.part_dqk = None
        self.part_db = None

    def ensure_bwd(self, bag):
        if self.dz is None:
            dev = bag.x.device
            self.dz = torch.empty((bag.total_rows, D), dtype=torch.bfloat16, device=dev)
            self.part_dqk = torch.empty((bag.num_tiles, Q, D), dtype=torch.float32, device=dev)
            self.part_db = torch.empty((bag.num_tiles, D), dtype=torch.float32, device=dev)
            if self.nacagat:
                self.dkg = torch.empty((bag.total_rows, D), dtype=torch.float16, device=dev)
                self.dg = torch.empty((Q, bag.total_rows), dtype=torch.float32, device=dev)
                self.part_dtq = torch.empty((bag.num_tiles, Q, D), dtype=torch.float32, device=dev)
                self.part_dbk = torch.empty((bag.num_tiles, D), dtype=torch.float32, device=dev)
                self.part_dkc = torch.empty((bag.num_tiles, 8), dtype=torch.float32, device=dev)
                self.dg_max = torch.zeros(1, dtype=torch.int32, device=dev)


def bag_forward(bag, w_h_bf16, bias_h, qk, ws, seed=0, drop_p=0.0, seed_dev=None):
    """mpo_bag_fwd: fills ws.scores / ws.pooled / ws.lse (and ws.h_saved when allocated)."""
    _lib.call("mpo_bag_fwd", bag.c(), _ptr(w_h_bf16), _ptr(bias_h), _ptr(qk), _ptr(ws.scores), _ptr(ws.part_ml),
              _ptr(ws.part_pool), _ptr(ws.pooled), _ptr(ws.lse), _ptr(ws.h_saved), None, ctypes.c_uint32(seed & 0xFFFFFFFF),
              _ptr(seed_dev), ctypes.c_float(drop_p), _stream())


def attention_map(bag, ws, out=None, seed=0, seed_dev=None, attn_drop_p=0.0):
    if out is None:
        out = torch.empty_like(ws.scores)
    if ws.nacagat:
        _lib.call("mpo_attn_map_dropout", bag.c(), _ptr(ws.scores_g), _ptr(ws.lse), _ptr(out),
                  ctypes.c_uint32(seed & 0xFFFFFFFF), _ptr(seed_dev), ctypes.c_float(attn_drop_p), _stream())
    else:
        _lib.call("mpo_attn_map", bag.c(), _ptr(ws.scores), _ptr(ws.lse), _ptr(out), _stream())
    return out


def cast_f16(src, out=None):
    """fp32 -> fp16 through mpo_cast_f16."""
    require_cuda(src, "cast_f16 input")
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, dtype=torch.float16, device=src.device)
    _lib.call("mpo_cast_f16", _ptr(src), _ptr(out), src.numel(), _stream())
    return out


def bag_project(bag, w_h_bf16, bias_h, qk, ws, seed=0, drop_p=0.0, seed_dev=None):
    """mpo_bag_fwd with pooled == NULL: activations (ws.h_saved) and raw folded scores only (NaCAGaT)."""
    _lib.call("mpo_bag_fwd", bag.c(), _ptr(w_h_bf16), _ptr(bias_h), _ptr(qk), _ptr(ws.scores), None, None, None, None,
              _ptr(ws.h_saved), _ptr(ws.h_lo), ctypes.c_uint32(seed & 0xFFFFFFFF), _ptr(seed_dev), ctypes.c_float(drop_p),
              _stream())


def bag_gate_forward(bag, w_k_f16, bias_k, qp, kc, ws, seed=0, attn_drop_p=0.0, seed_dev=None):
    """mpo_bag_gate_fwd: fills ws.scores_g / ws.pooled / ws.lse / ws.suma (and ws.pgate, ws.t_saved when allocated)."""
    _lib.call("mpo_bag_gate_fwd", bag.c(), _ptr(ws.h_saved), _ptr(ws.h_lo), _ptr(w_k_f16), _ptr(bias_k), _ptr(qp), _ptr(kc),
              _ptr(ws.scores), _ptr(ws.scores_g), _ptr(ws.pgate), _ptr(ws.t_saved), _ptr(ws.part_ml), _ptr(ws.part_pool),
              _ptr(ws.pooled), _ptr(ws.lse), _ptr(ws.suma), _ptr(ws.part_pool_lo), _ptr(ws.pooled_lo),
              ctypes.c_uint32(seed & 0xFFFFFFFF), _ptr(seed_dev),
              ctypes.c_float(attn_drop_p), _stream())


def attention_map_dot(bag, amap, other):
    """mpo_attn_map_dot: [B, 6] per-slide, per-query sums of amap * other over the patches."""
    out = torch.empty((bag.num_slides, Q), dtype=torch.float32, device=bag.x.device)
    _lib.call("mpo_attn_map_dot", bag.c(), _ptr(amap), _ptr(other), _ptr(out), _stream())
    return out


def bag_backward(bag, ws, dpooled, qk, grad_w_h, grad_b_h, drop_p=0.0, d_amap=None, amap_dot=None):
    """mpo_bag_bwd: accumulates into grad_w_h [256,1024] / grad_b_h [256], returns dqk [B,6,256]."""
    if ws.h_saved is None:
        raise RuntimeError("bag_backward needs the activations saved by bag_forward (save_h=True)")
    ws.ensure_bwd(bag)
    dqk = torch.empty((bag.num_slides, Q, D), dtype=torch.float32, device=bag.x.device)
    _lib.call("mpo_bag_bwd", bag.c(), _ptr(ws.h_saved), _ptr(ws.scores), _ptr(ws.lse), _ptr(ws.pooled),
              _ptr(dpooled), _ptr(qk), _ptr(ws.dz), _ptr(ws.part_dqk), _ptr(ws.part_db), _ptr(dqk),
              _ptr(grad_w_h), _ptr(grad_b_h), _ptr(d_amap), _ptr(amap_dot), ctypes.c_float(drop_p), _stream())
    return dqk


def lse_combine(lse_all, pooled_all):
    """Combine per-shard (lse [S,6], pooled [S,6,256]) of one patch-range-sharded bag."""
    S = lse_all.shape[0]
    lse = torch.empty((Q,), dtype=torch.float32, device=lse_all.device)
    pooled = torch.empty((Q, D), dtype=torch.float32, device=lse_all.device)
    _lib.call("mpo_lse_combine", _ptr(lse_all.contiguous()), _ptr(pooled_all.contiguous()), S, _ptr(lse),
              _ptr(pooled), _stream())
    return lse, pooled
