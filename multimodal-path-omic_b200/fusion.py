"""Fusion-layer parameter containers (reference: models/fusion.py).

Inside MCAT / NaCAGaT the fusion layers run in the slide tail (csrc/tail.cu / tail_fused.cu, stages "fusion"); these
classes hold the parameters under the reference's names and initialise them the same way.  Called on their own, as the
reference's unit tests do (models/fusion.py:116-170), they run operator by operator on the same CUDA kernels (ops.py:
inference only, no autograd graph, CUDA tensors only)."""
import torch
import torch.nn as nn

from .utils import init_max_weights


class ConcatFusion(nn.Module):
    """concat -> Linear -> ReLU -> Linear -> ReLU -- reference: models/fusion.py:7-19."""

    def __init__(self, dims: list, hidden_size: int = 256, output_size: int = 256):
        super().__init__()
        self.fusion_layer = nn.Sequential(nn.Linear(sum(dims), hidden_size), nn.ReLU(),
                                          nn.Linear(hidden_size, output_size), nn.ReLU())

    def forward(self, *x):
        """fusion_layer(cat(x)) -- models/fusion.py:17-19 (1-D inputs, as the reference's drivers pass them)."""
        from . import ops
        l0, l2 = self.fusion_layer[0], self.fusion_layer[2]
        z = ops.linear(torch.cat(x, dim=0), l0.weight, l0.bias, act="relu")
        return ops.linear(z, l2.weight, l2.bias, act="relu")


class GatedConcatFusion(nn.Module):
    """reference: models/fusion.py:22-41.  Its per-input gates live in a plain Python list there, so they are not
    parameters, never trained and never moved to the device; the state_dict only has fusion_layer.{0,2}.*.
    The gates are kept here the same way; the slide tail applies them from device copies (csrc/tail_kernels.cuh,
    gate_concat_*_kernel)."""

    def __init__(self, dims: list, hidden_size: int = 256, output_size: int = 256):
        super().__init__()
        self.gates = [nn.Sequential(nn.Linear(dim, 1), nn.Sigmoid()) for dim in dims]
        self.fusion_layer = nn.Sequential(nn.Linear(sum(dims), hidden_size), nn.ReLU(),
                                          nn.Linear(hidden_size, output_size), nn.ReLU())

    def forward(self, *x):
        """each input scaled by its own sigmoid gate, then concat + MLP -- models/fusion.py:34-41.  The gates are the
        plain-list modules of the reference (never registered, so `.to(device)` does not move them): their weights are
        copied to the inputs' device for the call."""
        from . import ops
        items = []
        for gate, item in zip(self.gates, x):
            lin = gate[0]
            g = ops.linear(item, lin.weight.to(item.device), lin.bias.to(item.device), act="sigmoid")
            items.append(ops.rowscale(item, g))
        l0, l2 = self.fusion_layer[0], self.fusion_layer[2]
        z = ops.linear(torch.cat(items, dim=0), l0.weight, l0.bias, act="relu")
        return ops.linear(z, l2.weight, l2.bias, act="relu")


class BilinearFusion(nn.Module):
    """Gated bilinear (Kronecker) fusion -- reference: models/fusion.py:44-113."""

    def __init__(self, dim1: int = 256, dim2: int = 256, hidden_size: int = 32, output_size: int = 64,
                 mm_hidden_size: int = 64, use_skip_connection=True, use_bilinear=True, use_gates=True, dropout=0.25):
        super().__init__()
        self.use_skip_connection = use_skip_connection
        self.use_bilinear = use_bilinear
        self.use_gates = use_gates
        self.dropout = dropout

        def gate_branch(d_a, d_b):
            lin_h = nn.Sequential(nn.Linear(d_a, hidden_size), nn.ReLU())
            lin_z = nn.Bilinear(d_a, d_b, hidden_size) if use_bilinear else nn.Linear(d_a + d_b, hidden_size)
            lin_o = nn.Sequential(nn.Linear(hidden_size, hidden_size), nn.ReLU(), nn.Dropout(p=dropout))
            return lin_h, lin_z, lin_o

        self.linear_h1, self.linear_z1, self.linear_o1 = gate_branch(dim1, dim2)
        self.linear_h2, self.linear_z2, self.linear_o2 = gate_branch(dim2, dim1)
        self.post_fusion_dropout = nn.Dropout(p=dropout)
        self.fc1 = nn.Sequential(nn.Linear((hidden_size + 1) * (hidden_size + 1), mm_hidden_size), nn.ReLU(),
                                 nn.Dropout(p=dropout))
        self.fc2 = nn.Sequential(nn.Linear(mm_hidden_size + 2 * hidden_size + 2, output_size), nn.ReLU(),
                                 nn.Dropout(p=dropout))
        init_max_weights(self)

    def forward(self, *x):
        if len(x) != 2:
            raise RuntimeError('Bilinear fusion is possible only on 2 inputs')
        return self._forward_standalone(x[0], x[1])

    def _forward_standalone(self, x1, x2):
        """models/fusion.py:84-113 on the kernels of the slide tail's bilinear stage (csrc/tail_kernels.cuh: bil_gate /
        bil_kron), which are built for the reference's own configuration: 256-wide inputs, hidden 32, gated + bilinear."""
        import ctypes
        from . import _lib, ops
        from .bagpass import _ptr, _stream
        h1w = self.linear_h1[0].weight
        if not (self.use_gates and self.use_bilinear) or tuple(h1w.shape) != (32, 256) or \
                tuple(self.linear_h2[0].weight.shape) != (32, 256):
            raise NotImplementedError("stand-alone BilinearFusion runs the reference's default configuration "
                                      "(dim1 = dim2 = 256, hidden_size = 32, use_gates, use_bilinear)")
        p = float(self.dropout) if self.training else 0.0
        x1r, x2r = ops._prep(x1, "x1").reshape(1, -1), ops._prep(x2, "x2").reshape(1, -1)
        dev = x1r.device
        f32 = dict(dtype=torch.float32, device=dev)

        def side(xa, xb, lin_h, lin_z, lin_o):
            h = ops.linear(xa, lin_h[0].weight, lin_h[0].bias, act="relu")
            U = ops.linear(xb, lin_z.weight.reshape(32 * 256, 256))            # U[k * 256 + i] = sum_j W[k][i][j] xb[j]
            g, gh = torch.empty((1, 32), **f32), torch.empty((1, 32), **f32)
            _lib.call("mpo_op_bil_gate", _ptr(xa), _ptr(U), _ptr(ops._prep(lin_z.bias, "bias")), _ptr(h), _ptr(g), _ptr(gh),
                      1, _stream())
            return ops.linear(gh, lin_o[0].weight, lin_o[0].bias, act="relu", drop_p=p)

        o1 = side(x1r, x2r, self.linear_h1, self.linear_z1, self.linear_o1)
        o2 = side(x2r, x1r, self.linear_h2, self.linear_z2, self.linear_o2)
        mm = self.fc1[0].weight.shape[0]
        if mm != 64 and self.use_skip_connection:
            raise NotImplementedError("stand-alone BilinearFusion: mm_hidden_size = 64 with the skip connection")
        kp = torch.empty((1, 33 * 33), **f32)
        cat = torch.empty((1, 130), **f32)
        _lib.call("mpo_op_bil_kron", _ptr(o1), _ptr(o2), _ptr(kp), _ptr(cat), 1, ctypes.c_float(p),
                  ops._seed() if p > 0 else 0, next(ops._site) & 0xFFFF, _stream())
        if self.use_skip_connection:
            ops.linear_into(kp, self.fc1[0].weight, self.fc1[0].bias, cat, 130, act="relu", drop_p=p)   # cat[:, :64]
            z = cat
        else:
            z = ops.linear(kp, self.fc1[0].weight, self.fc1[0].bias, act="relu", drop_p=p)
        return ops.linear(z, self.fc2[0].weight, self.fc2[0].bias, act="relu", drop_p=p).reshape(-1)
