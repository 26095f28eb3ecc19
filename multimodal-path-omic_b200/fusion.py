"""Fusion-layer parameter containers (reference: models/fusion.py).

The arithmetic of ConcatFusion and BilinearFusion runs inside the slide tail (csrc/tail.cu, stages "fusion");
these classes hold the parameters under the reference's names and initialise them the same way."""
import torch.nn as nn

from .blocks import _standalone
from .utils import init_max_weights


class ConcatFusion(nn.Module):
    """concat -> Linear -> ReLU -> Linear -> ReLU -- reference: models/fusion.py:7-19."""

    def __init__(self, dims: list, hidden_size: int = 256, output_size: int = 256):
        super().__init__()
        self.fusion_layer = nn.Sequential(nn.Linear(sum(dims), hidden_size), nn.ReLU(),
                                          nn.Linear(hidden_size, output_size), nn.ReLU())

    def forward(self, *x):
        _standalone("ConcatFusion")


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
        _standalone("GatedConcatFusion")


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
        _standalone("BilinearFusion")
