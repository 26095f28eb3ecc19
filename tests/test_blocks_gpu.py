"""Stand-alone forward of the block / fusion classes (the reference's own unit tests call them directly:
models/blocks.py:304-325 test_cag, models/fusion.py:116-170 test_concat_fusion / test_gated_concat_fusion /
test_bilinear_fusion).  Same calls and shape checks as those tests; in addition the values are compared with the
reference arithmetic (models/blocks.py:42-48, 247-253; models/fusion.py:17-19, 34-41, 84-113) restated with torch
float64 on the CPU from the same parameters -- the checker, not the product: the product path is mpo_op_* kernels."""
import os
import sys
from importlib import import_module

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu
TOL = 1e-4          # fp32 kernels against float64: relative to the largest output


def _pkg(name):
    return import_module("multimodal-path-omic_b200." + name)


def _close(got, want):
    want = want.to(torch.float64)
    err = float((got.detach().cpu().to(torch.float64) - want).abs().max() / want.abs().max().clamp_min(1e-6))
    assert err < TOL, err


def _lin(layer, x):
    return x @ layer.weight.detach().cpu().double().t() + layer.bias.detach().cpu().double()


def test_concat_fusion_as_the_reference_test_calls_it():
    fusion = _pkg("fusion")
    torch.manual_seed(0)
    for n_in, kw, out_len in ((2, {}, 256), (3, dict(hidden_size=128, output_size=32), 32)):
        xs = [torch.randn(256) for _ in range(n_in)]
        m = fusion.ConcatFusion(dims=[256] * n_in, **kw).cuda().eval()
        out = m(*[x.cuda() for x in xs])
        assert len(out) == out_len and out.is_cuda
        cat = torch.cat(xs).double()
        want = torch.relu(_lin(m.fusion_layer[2], torch.relu(_lin(m.fusion_layer[0], cat))))
        _close(out, want)


def test_gated_concat_fusion_as_the_reference_test_calls_it():
    fusion = _pkg("fusion")
    torch.manual_seed(1)
    for n_in, kw, out_len in ((2, {}, 256), (3, dict(hidden_size=128, output_size=32), 32)):
        xs = [torch.randn(256) for _ in range(n_in)]
        m = fusion.GatedConcatFusion(dims=[256] * n_in, **kw).cuda().eval()      # the gates stay on the CPU (plain list)
        out = m(*[x.cuda() for x in xs])
        assert len(out) == out_len
        items = [x.double() * torch.sigmoid(_lin(g[0], x.double())) for g, x in zip(m.gates, xs)]
        want = torch.relu(_lin(m.fusion_layer[2], torch.relu(_lin(m.fusion_layer[0], torch.cat(items)))))
        _close(out, want)


def test_bilinear_fusion_as_the_reference_test_calls_it():
    fusion = _pkg("fusion")
    torch.manual_seed(2)
    for kw, out_len in (({}, 64), (dict(output_size=256), 256)):
        x1, x2 = torch.randn(256), torch.randn(256)
        m = fusion.BilinearFusion(dim1=256, dim2=256, **kw).cuda().eval()
        out = m(x1.cuda(), x2.cuda())
        assert len(out) == out_len

        def side(xa, xb, lh, lz, lo):
            h = torch.relu(_lin(lh[0], xa))
            z = torch.einsum("i,kij,j->k", xa, lz.weight.detach().cpu().double(), xb) + lz.bias.detach().cpu().double()
            return torch.relu(_lin(lo[0], torch.sigmoid(z) * h))
        a, b = x1.double(), x2.double()
        o1 = torch.cat([side(a, b, m.linear_h1, m.linear_z1, m.linear_o1), torch.ones(1, dtype=torch.float64)])
        o2 = torch.cat([side(b, a, m.linear_h2, m.linear_z2, m.linear_o2), torch.ones(1, dtype=torch.float64)])
        kp = torch.outer(o1, o2).flatten()
        f1 = torch.relu(_lin(m.fc1[0], kp))
        want = torch.relu(_lin(m.fc2[0], torch.cat([f1, o1, o2])))
        _close(out, want)
    with pytest.raises(RuntimeError):
        m(x1.cuda())                                        # "Bilinear fusion is possible only on 2 inputs"


def test_contextual_attention_gate_as_the_reference_test_calls_it():
    blocks = _pkg("blocks")
    torch.manual_seed(3)
    x1, x2 = torch.randn(8, 256), torch.randn(8, 256)
    for hidden in (256, 128):
        m = blocks.ContextualAttentionGate(hidden_dim=hidden).cuda().eval()
        C = m(x1.cuda(), x2.cuda())
        assert C.shape[0] == 8 and C.shape[1] == hidden
        elu = torch.nn.functional.elu

        def ln(seq, v):
            n = seq[1]
            return torch.nn.functional.layer_norm(elu(v), (hidden,), n.weight.detach().cpu().double(),
                                                  n.bias.detach().cpu().double(), n.eps)
        a, b = x1.double(), x2.double()
        G = ln(m.G, elu(_lin(m.fc1[0], a)) + elu(_lin(m.fc2[0], b)))
        Eg = ln(m.E, elu(_lin(m.fc3[0], b)))
        _close(C, elu(_lin(m.fc_c[0], G * Eg)))


def test_attention_net_gated_forward_and_train_mode_dropout():
    blocks = _pkg("blocks")
    torch.manual_seed(4)
    x = torch.randn(6, 256)
    m = blocks.AttentionNetGated(n_classes=1).cuda().eval()
    A, x_out = m(x.cuda())
    assert tuple(A.shape) == (6, 1) and x_out.shape == x.shape
    a = torch.tanh(_lin(m.attention_a[0], x.double()))
    b = torch.sigmoid(_lin(m.attention_b[0], x.double()))
    _close(A, _lin(m.attention_c, a * b))
    m.train()                                              # p = 0.25 on both branches: the output changes, stays finite
    A2, _ = m(x.cuda())
    assert torch.isfinite(A2).all() and not torch.allclose(A2, A)


def test_standalone_forward_refuses_cpu_tensors():
    fusion = _pkg("fusion")
    m = fusion.ConcatFusion(dims=[256, 256])
    with pytest.raises(RuntimeError):
        m(torch.randn(256), torch.randn(256))
