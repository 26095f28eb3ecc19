"""Peer-memory groups for the hand-written NVLink collectives of csrc/peer.cu (one process per GPU).

torch.distributed is used ONCE, at set-up, to exchange the CUDA-IPC handles of the exchange / gradient / parameter
buffers; afterwards every exchange of the hot path is a kernel launch on the caller's stream (capturable in CUDA graphs).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .bagpass import D, Q, _ptr, _stream


class _RawCuda:
    """__cuda_array_interface__ view of library-owned device memory (so that torch can address it without owning it)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class PeerBuffer:
    """Device memory from mpo_peer_alloc: its own cudaMalloc allocation, hence exportable over CUDA IPC at offset 0."""

    def __init__(self, nbytes, device):
        self.nbytes = int(nbytes)
        self.device = torch.device(device)
        p = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.call("mpo_peer_alloc", self.nbytes, ctypes.byref(p))
        self.ptr = int(p.value)
        self._raw = _RawCuda(self.ptr, self.nbytes)
        self.bytes = torch.as_tensor(self._raw, device=self.device)

    def tensor(self, dtype, shape=None):
        t = self.bytes.view(dtype)
        return t if shape is None else t[:int(torch.Size(shape).numel())].view(shape)

    def export(self):
        h = ctypes.create_string_buffer(64)
        _lib.call("mpo_peer_export", ctypes.c_void_p(self.ptr), h)
        return bytes(h.raw)


class PeerGroup:
    def __init__(self, device, group=None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerGroup needs an initialised torch.distributed process group (for the handle exchange)")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("peer collectives are built for one NVSwitch node (<= 8 GPUs)")
        self.device = torch.device(device)
        self.exchange = PeerBuffer(_lib.lib().mpo_peer_exchange_bytes(), self.device)
        self.epochs = torch.zeros(8, dtype=torch.int32, device=self.device)
        self._opened = []
        self.c = _lib.MpoPeerGroup()
        self.c.world, self.c.rank = self.world, self.rank
        self.c.epochs = self.epochs.data_ptr()
        off = _lib.lib().mpo_peer_flags_offset()
        for p, ptr in enumerate(self.share(self.exchange)):
            self.c.data[p] = ptr
            self.c.flags[p] = ptr + off
        with torch.cuda.device(self.device):
            _lib.call("mpo_peer_warmup")

    def share(self, buf):
        """exports `buf`, gathers every rank's handle and maps the others: -> list of W device pointers valid HERE."""
        handles = [None] * self.world
        dist.all_gather_object(handles, buf.export(), group=self.group)
        ptrs = []
        with torch.cuda.device(self.device):
            for p, h in enumerate(handles):
                if p == self.rank:
                    ptrs.append(buf.ptr)
                    continue
                out = ctypes.c_void_p()
                _lib.call("mpo_peer_open", ctypes.create_string_buffer(h, 64), ctypes.byref(out))
                self._opened.append(int(out.value))
                ptrs.append(int(out.value))
        return ptrs

    def ref(self):
        return ctypes.byref(self.c)

    def barrier(self, slot=7):
        _lib.call("mpo_peer_barrier", self.ref(), slot, _stream())

    def lse_combine(self, lse_local, pooled_local, lse_out=None, pooled_out=None, slot=0):
        """mpo_peer_lse_combine: (lse [6], pooled [6,256]) of this rank's patch range -> the merged state of the bag."""
        if lse_out is None:
            lse_out = torch.empty(Q, dtype=torch.float32, device=self.device)
            pooled_out = torch.empty((Q, D), dtype=torch.float32, device=self.device)
        _lib.call("mpo_peer_lse_combine", self.ref(), slot, _ptr(lse_local), _ptr(pooled_local), _ptr(lse_out),
                  _ptr(pooled_out), _stream())
        return lse_out, pooled_out

    def register_flat_buffers(self, grad_buf, param_buf):
        for p, ptr in enumerate(self.share(grad_buf)):
            self.c.grad[p] = ptr
        for p, ptr in enumerate(self.share(param_buf)):
            self.c.param[p] = ptr

    def slice_of(self, lo, hi, rank=None):
        a, b = ctypes.c_int64(), ctypes.c_int64()
        _lib.lib().mpo_peer_slice(lo, hi, self.world, self.rank if rank is None else rank, ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    def close(self):
        """unmap the peers' buffers (call on every rank before the process group is destroyed)."""
        torch.cuda.synchronize(self.device)
        for ptr in self._opened:
            _lib.lib().mpo_peer_close(ctypes.c_void_p(ptr))
        self._opened = []
