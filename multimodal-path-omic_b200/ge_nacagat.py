"""Drop-in GeneExprNarrowContextualAttentionGateTransformer (reference: models/ge_nacagat/ge_nacagat.py).

Same constructor / forward(wsi) / outputs / state_dict keys as ge_nacagat.py:9-72: Y [n_classes] (soft-maxed), and
attention_scores = {'attn': [N, N] self-attention map, 'path': [1, N] pooling logits}.  The device work runs through the
C ABI (mpo_bag_fwd projection pass, mpo_ge_fwd / mpo_ge_bwd); there is no CPU fallback."""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from . import bagpass as bp
from .bagpass import D, Q, _ptr, _stream, require_cuda
from .blocks import AttentionNetGated
from .slidepath import _next_seed

_WIDTH = {"small": 128, "medium": 256, "big": 512}


class GeneExprNarrowContextualAttentionGateTransformer(nn.Module):
    def __init__(self, model_size: str = 'medium', n_classes: int = 3, dropout: float = 0.25):
        super().__init__()
        if model_size in _WIDTH:
            self.model_sizes = [_WIDTH[model_size], _WIDTH[model_size]]
        w0, w1 = self.model_sizes          # AttributeError for an unknown size, as in the reference
        self.n_classes = n_classes
        self.dropout = dropout
        # creation order = the reference's (ge_nacagat.py:20-39), so that seeds give identical initial weights
        self.H = nn.Sequential(nn.Linear(1024, w0), nn.ReLU(), nn.Dropout(dropout))
        self.self_attention = nn.MultiheadAttention(embed_dim=w1, num_heads=1)
        layer = nn.TransformerEncoderLayer(d_model=w1, nhead=8, dim_feedforward=512, dropout=dropout, activation='relu')
        self.path_transformer = nn.TransformerEncoder(layer, num_layers=2)
        self.path_attention_head = AttentionNetGated(n_classes=1, input_dim=w1, hidden_dim=w1)
        self.path_rho = nn.Sequential(*[nn.Linear(w1, w1), nn.ReLU(), nn.Dropout(dropout)])
        self.classifier = nn.Linear(w1, n_classes)
        self._w_bf16 = None
        self._master_ref = (self,)      # nn.DataParallel replicas have empty _parameters: they run on this module's leaves

    def get_trainable_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    # -- struct mpo_ge_model over this module's parameters
    def _binding(self, grads=None):
        if self.model_sizes[0] != D:
            raise NotImplementedError("the B200 kernels are built for model_size='medium' (256)")
        P = dict(self.named_parameters())
        for n, p in P.items():
            if not p.is_cuda:
                raise RuntimeError("parameter %s is on %s: move the model to a CUDA device (no CPU fallback)" % (n, p.device))

        def lin(w, b):
            L = _lib.MpoLin()
            L.w, L.b = P[w].data_ptr(), P[b].data_ptr()
            if grads is not None:
                L.gw, L.gb = grads[w].data_ptr(), grads[b].data_ptr()
            return L

        def norm(prefix):
            Nn = _lib.MpoNorm()
            Nn.g, Nn.b = P[prefix + ".weight"].data_ptr(), P[prefix + ".bias"].data_ptr()
            if grads is not None:
                Nn.gg, Nn.gb = grads[prefix + ".weight"].data_ptr(), grads[prefix + ".bias"].data_ptr()
            return Nn

        m = _lib.MpoGeModel()
        m.n_classes = self.n_classes
        m.H = lin("H.0.weight", "H.0.bias")
        m.sa_in = lin("self_attention.in_proj_weight", "self_attention.in_proj_bias")
        m.sa_out = lin("self_attention.out_proj.weight", "self_attention.out_proj.bias")
        for l in range(2):
            pre = "path_transformer.layers.%d" % l
            E = _lib.MpoEncoderLayer()
            E.in_proj = lin(pre + ".self_attn.in_proj_weight", pre + ".self_attn.in_proj_bias")
            E.out_proj = lin(pre + ".self_attn.out_proj.weight", pre + ".self_attn.out_proj.bias")
            E.linear1 = lin(pre + ".linear1.weight", pre + ".linear1.bias")
            E.linear2 = lin(pre + ".linear2.weight", pre + ".linear2.bias")
            E.norm1, E.norm2 = norm(pre + ".norm1"), norm(pre + ".norm2")
            m.tr[l] = E
        H = _lib.MpoPoolHead()
        H.att_a = lin("path_attention_head.attention_a.0.weight", "path_attention_head.attention_a.0.bias")
        H.att_b = lin("path_attention_head.attention_b.0.weight", "path_attention_head.attention_b.0.bias")
        H.att_c = lin("path_attention_head.attention_c.weight", "path_attention_head.attention_c.bias")
        H.rho = lin("path_rho.0.weight", "path_rho.0.bias")
        m.pool = H
        m.classifier = lin("classifier.weight", "classifier.bias")
        m._keep = (P, grads)
        return m

    def forward(self, wsi):
        require_cuda(wsi, "wsi")
        master = self._master_ref[0]
        names = [n for n, _ in master.named_parameters()]
        params = [p for _, p in master.named_parameters()]
        needs_bwd = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        Y, attn, path = _GeFn.apply(master, needs_bwd, bool(self.training), wsi, len(names), *params)
        return Y, {'attn': attn, 'path': path}


class _GeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, needs_bwd, train, wsi, n_params, *params):
        bag = bp.PackedBag.from_slides([wsi.detach()])
        N = bag.total_rows
        dev = bag.x.device
        model = module._binding(grads=None)
        w_h = dict(module.named_parameters())["H.0.weight"]
        if module._w_bf16 is None or module._w_bf16.device != dev:
            module._w_bf16 = torch.empty((D, 1024), dtype=torch.bfloat16, device=dev)
        bp.cast_bf16(w_h.detach(), out=module._w_bf16)
        ws_bag = bp.BagWorkspace(bag, save_h=True, nacagat=True, save_gate=False)
        qk0 = torch.zeros((1, Q, D), dtype=torch.float32, device=dev)
        drop_p = float(module.dropout) if train else 0.0
        seed = (getattr(module, "_fixed_seed", None) or _next_seed()) if train else 0
        bp.bag_project(bag, module._w_bf16, dict(module.named_parameters())["H.0.bias"].detach(), qk0, ws_bag,
                       seed=seed, drop_p=drop_p)
        nfl = _lib.lib().mpo_ge_ws_floats(N)
        if nfl <= 0:
            raise RuntimeError("mpo_ge_ws_floats failed")
        ws = torch.empty(nfl, dtype=torch.float32, device=dev)
        attn = torch.empty((N, N), dtype=torch.float32, device=dev)
        path = torch.empty((1, N), dtype=torch.float32, device=dev)
        Y = torch.empty(module.n_classes, dtype=torch.float32, device=dev)
        _lib.call("mpo_ge_fwd", ctypes.byref(model), N, _ptr(ws_bag.h_saved), _ptr(ws_bag.h_lo), _ptr(ws), _ptr(attn),
                  _ptr(path), _ptr(Y), ctypes.c_float(drop_p), ctypes.c_uint32(seed & 0xFFFFFFFF), 1 if train else 0, _stream())
        if needs_bwd:
            ctx.saved = (module, bag, ws_bag, ws, attn, path, Y, drop_p, seed, train)
        else:
            ctx.saved = None
            del ws
        ctx.n_params = n_params
        ctx.mark_non_differentiable(attn, path)
        return Y, attn, path

    @staticmethod
    def backward(ctx, dY, *unused):
        if ctx.saved is None:
            raise RuntimeError("this forward pass did not keep activations (it ran under no_grad)")
        module, bag, ws_bag, ws, attn, path, Y, drop_p, seed, train = ctx.saved
        dev = bag.x.device
        P = dict(module.named_parameters())
        offs, off = {}, 0
        for n, p in P.items():
            offs[n] = off
            off += (p.numel() + 63) // 64 * 64
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        grads = {n: flat[offs[n]:offs[n] + p.numel()].view_as(p) for n, p in P.items()}
        model = module._binding(grads=grads)
        dz = torch.empty((bag.total_rows, D), dtype=torch.bfloat16, device=dev)
        keep = 1.0 / (1.0 - drop_p) if drop_p > 0 else 1.0
        dYc = dY.detach().to(torch.float32).contiguous()
        _lib.call("mpo_ge_bwd", ctypes.byref(model), bag.c(), _ptr(ws_bag.h_saved), _ptr(ws), _ptr(attn), _ptr(path),
                  _ptr(Y), _ptr(dYc), _ptr(dz), ctypes.c_float(keep), ctypes.c_float(drop_p),
                  ctypes.c_uint32(seed & 0xFFFFFFFF), 1 if train else 0, _stream())
        return (None, None, None, None, None) + tuple(grads[n] for n in P.keys())


def ge_cross_entropy(Y, label):
    """The reference driver's loss (models/ge_nacagat/main.py:29,33): nn.CrossEntropyLoss()(Y.unsqueeze(0), label) on the
    already soft-maxed Y, computed by mpo_ge_ce_loss.  Returns a scalar tensor with autograd support."""
    return _GeLossFn.apply(Y, label)


class _GeLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Y, label):
        require_cuda(Y, "Y")
        Yc = Y.detach().to(torch.float32).contiguous()
        lab = label.detach().to(device=Y.device, dtype=torch.int64).reshape(-1).contiguous()
        loss = torch.empty(1, dtype=torch.float32, device=Y.device)
        dY = torch.empty_like(Yc)
        _lib.call("mpo_ge_ce_loss", _ptr(Yc), _ptr(lab), Yc.numel(), ctypes.c_float(1.0), _ptr(loss), _ptr(dY), _stream())
        ctx.save_for_backward(dY)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dY,) = ctx.saved_tensors
        return dY * g, None
