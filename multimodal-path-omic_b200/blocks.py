"""Parameter containers mirroring the reference's building blocks (models/blocks.py).

These classes own the nn.Parameters under the reference's state_dict names and reproduce its initialisation
(same torch initialisers called in the same order, so a reference checkpoint -- or the same torch seed -- gives
identical weights).  They carry no PyTorch arithmetic: inside MCAT / NaCAGaT the slide engine reads the parameter
storage through the C ABI and runs them fused into the slide pass.  Called on their own (as the reference's unit tests
do, models/blocks.py:288-345), AttentionNetGated and ContextualAttentionGate run operator by operator on the same
CUDA kernels (ops.py: inference only, no autograd graph); PreGatingContextualAttention exists only fused into the bag
pass and raises.
"""
import torch
import torch.nn as nn
from torch.nn.init import constant_, xavier_uniform_
from torch.nn.modules.linear import NonDynamicallyQuantizableLinear


def _standalone(name):
    raise RuntimeError(
        "%s has no stand-alone kernel on the B200 path: it runs fused inside the MCAT / NaCAGaT slide pass "
        "(multimodal-path-omic_b200.slidepath). There is no PyTorch fallback." % name)


class AttentionNetGated(nn.Module):
    """Gated attention scorer -- reference: models/blocks.py:13-48.

    State: attention_a.0 / attention_b.0 (Linear input_dim -> hidden_dim behind Tanh / Sigmoid) and attention_c
    (Linear hidden_dim -> n_classes).  The reference hard-codes p = 0.25 dropout on both branches when dropout_p
    is truthy (blocks.py:34-36); the engine applies the same rate in train mode."""

    def __init__(self, input_dim: int = 256, hidden_dim: int = 256, dropout_p: bool = True, n_classes: int = 1):
        super().__init__()
        branch_a = [nn.Linear(input_dim, hidden_dim), nn.Tanh()]
        branch_b = [nn.Linear(input_dim, hidden_dim), nn.Sigmoid()]
        if dropout_p:
            branch_a.append(nn.Dropout(0.25))
            branch_b.append(nn.Dropout(0.25))
        self.attention_a = nn.Sequential(*branch_a)
        self.attention_b = nn.Sequential(*branch_b)
        self.attention_c = nn.Linear(hidden_dim, n_classes)
        self.branch_dropout = 0.25 if dropout_p else 0.0

    def forward(self, x):
        """(A, x) as models/blocks.py:42-48: A = attention_c(tanh-branch * sigmoid-branch), both branches with p = 0.25
        dropout in train mode."""
        from . import ops
        p = self.branch_dropout if self.training else 0.0
        la, lb = self.attention_a[0], self.attention_b[0]
        a = ops.linear(x, la.weight, la.bias, act="tanh", drop_p=p)
        b = ops.linear(x, lb.weight, lb.bias, act="sigmoid", drop_p=p)
        A = ops.linear(ops.mul(a, b), self.attention_c.weight, self.attention_c.bias)
        return A, x


class ContextualAttentionGate(nn.Module):
    """Contextual attention gate -- reference: models/blocks.py:232-253."""

    def __init__(self, dim: int = 256, hidden_dim: int = 128):
        super().__init__()
        self.fc1 = nn.Sequential(nn.Linear(dim, hidden_dim), nn.ELU())
        self.fc2 = nn.Sequential(nn.Linear(dim, hidden_dim), nn.ELU())
        self.fc3 = nn.Sequential(nn.Linear(dim, hidden_dim), nn.ELU())
        self.G = nn.Sequential(nn.ELU(), nn.LayerNorm(hidden_dim))
        self.E = nn.Sequential(nn.ELU(), nn.LayerNorm(hidden_dim))
        self.fc_c = nn.Sequential(nn.Linear(hidden_dim, hidden_dim), nn.ELU())

    def forward(self, Q, Q_hat):
        """C = fc_c(G(fc1(Q) + fc2(Q_hat)) * E(fc3(Q_hat))) -- models/blocks.py:247-253."""
        from . import ops
        f1 = ops.linear(Q, self.fc1[0].weight, self.fc1[0].bias, act="elu")
        f2 = ops.linear(Q_hat, self.fc2[0].weight, self.fc2[0].bias, act="elu")
        f3 = ops.linear(Q_hat, self.fc3[0].weight, self.fc3[0].bias, act="elu")
        G = ops.layernorm(ops.act(ops.add(f1, f2), "elu"), self.G[1].weight, self.G[1].bias, self.G[1].eps)
        Eg = ops.layernorm(ops.act(f3, "elu"), self.E[1].weight, self.E[1].bias, self.E[1].eps)
        return ops.linear(ops.mul(G, Eg), self.fc_c[0].weight, self.fc_c[0].bias, act="elu")


class PreGatingContextualAttention(nn.Module):
    """NaCAGaT co-attention (tanh pre-gate + contextual gate) -- reference: models/blocks.py:51-111.

    Parameters: in_proj_weight [3E, E], in_proj_bias [3E], out_proj, CAG.*; xavier-uniform in_proj_weight and
    zero biases exactly as blocks.py:81-90."""

    def __init__(self, embed_dim, num_heads, device=None, dtype=None, dropout_p: float = 0.25) -> None:
        kw = {"device": device, "dtype": dtype}
        super().__init__()
        if embed_dim % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        self.embed_dim = self.kdim = self.vdim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.dropout = dropout_p
        self.batch_first = False
        self.in_proj_weight = nn.Parameter(torch.empty((3 * embed_dim, embed_dim), **kw))
        for unused in ("q_proj_weight", "k_proj_weight", "v_proj_weight"):
            self.register_parameter(unused, None)
        self.in_proj_bias = nn.Parameter(torch.empty(3 * embed_dim, **kw))
        self.out_proj = NonDynamicallyQuantizableLinear(embed_dim, embed_dim, bias=True, **kw)
        self.bias_k = self.bias_v = None
        self.add_zero_attn = False
        self.CAG = ContextualAttentionGate(dim=embed_dim, hidden_dim=embed_dim)
        xavier_uniform_(self.in_proj_weight)
        constant_(self.in_proj_bias, 0.0)
        constant_(self.out_proj.bias, 0.0)

    def forward(self, query, key, value, attn_mask=None, average_attn_weights=True, is_causal=False):
        _standalone("PreGatingContextualAttention")
