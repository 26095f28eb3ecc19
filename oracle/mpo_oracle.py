"""CPU oracle for the slide hot path -- TEST INFRASTRUCTURE ONLY.

A plain numpy (float64) restatement of what the reference's PyTorch modules compute for one slide, forward and
backward, in eval mode (dropout off).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this file; nothing under multimodal-path-omic_b200/ does.

Pinned against the reference: tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the
unmodified reference modules from /root/reference (outputs and autograd gradients); tests/test_oracle.py checks
this file against them and against the reference's only known-answer test (models/loss.py:104-121).

The arithmetic of nn.MultiheadAttention / nn.TransformerEncoderLayer lives in PyTorch itself (the reference has
no lockfile; the container has torch 2.11.0).  Their published algorithm is restated here and anchored on the
reference's call sites: models/mcat/mcat.py:48,51-53,60-62,97,101-102 and models/blocks.py:114-206.

Parameters are passed as a dict  state_dict-key -> numpy array  (same keys as the reference's state_dict).
"""
import numpy as np

F64 = np.float64      # working precision; bench.py's cpu_baseline leg switches it to float32 with use_dtype()
LN_EPS = 1e-5


def use_dtype(dtype):
    """float64 (default, the checker) or float32 (timing the reference algorithm at the reference's precision)."""
    global F64
    F64 = dtype


# ------------------------------------------------------------------------------------------------ primitives
def linear(x, W, b=None):
    y = x @ W.T
    return y if b is None else y + b


def linear_bwd(dy, x, W):
    """returns dx, dW, db for y = x W^T + b with x [R,in], dy [R,out]."""
    return dy @ W, dy.T @ x, dy.sum(axis=0)


def elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def elu_bwd(dy, y):
    return dy * np.where(y > 0, 1.0, y + 1.0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax(x, axis=-1):
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)


def softmax_bwd(da, a, axis=-1):
    return a * (da - (da * a).sum(axis=axis, keepdims=True))


def layernorm(x, g, b):
    mu = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + LN_EPS)
    xhat = (x - mu) * rstd
    return xhat * g + b, (xhat, rstd)


def layernorm_bwd(dy, cache, g):
    xhat, rstd = cache
    dxh = dy * g
    dx = rstd * (dxh - dxh.mean(axis=-1, keepdims=True) - xhat * (dxh * xhat).mean(axis=-1, keepdims=True))
    return dx, (dy * xhat).sum(axis=0), dy.sum(axis=0)


class Grads(dict):
    """parameter-gradient accumulator keyed like the state_dict."""

    def add(self, key, val):
        if key in self:
            self[key] = self[key] + val
        else:
            self[key] = np.array(val, dtype=F64)


# ------------------------------------------------------------------------------------------------ SNN omic encoders
# reference: models/mcat/mcat.py:32-45 (Linear+ELU+AlphaDropout twice), :90-92
def snn_fwd(P, omics):
    G, cache = [], []
    for i, x in enumerate(omics):
        x = np.asarray(x, F64).reshape(-1)
        g1 = elu(linear(x, P[f"G.{i}.0.0.weight"], P[f"G.{i}.0.0.bias"]))
        g2 = elu(linear(g1, P[f"G.{i}.1.0.weight"], P[f"G.{i}.1.0.bias"]))
        G.append(g2)
        cache.append((x, g1, g2))
    return np.stack(G), cache


def snn_bwd(P, cache, dG, grads):
    for i, (x, g1, g2) in enumerate(cache):
        d2 = elu_bwd(dG[i], g2)
        grads.add(f"G.{i}.1.0.weight", np.outer(d2, g1))
        grads.add(f"G.{i}.1.0.bias", d2)
        d1 = elu_bwd(d2 @ P[f"G.{i}.1.0.weight"], g1)
        grads.add(f"G.{i}.0.0.weight", np.outer(d1, x))
        grads.add(f"G.{i}.0.0.bias", d1)


# ------------------------------------------------------------------------------------------------ bag projection
# reference: models/mcat/mcat.py:24-29,87   H = Dropout(ReLU(Linear(1024,256)))
def bag_proj_fwd(P, X):
    H = np.maximum(linear(X, P["H.0.weight"], P["H.0.bias"]), 0.0)
    return H


def bag_proj_bwd(P, X, H, dH, grads):
    dZ = dH * (H > 0)
    grads.add("H.0.weight", dZ.T @ X)
    grads.add("H.0.bias", dZ.sum(axis=0))


# ------------------------------------------------------------------------------------------------ MCAT co-attention
# reference: models/mcat/mcat.py:48,97 -> torch.nn.functional.multi_head_attention_forward, one head, E = 256:
#   q = W_q G + b_q, k = W_k H + b_k, v = W_v H + b_v; a = softmax(q k^T / sqrt(E)); out = W_o (a v) + b_o
def mcat_coattn_fwd(P, G, H, pre="co_attention."):
    E = G.shape[1]
    Win, bin_ = P[pre + "in_proj_weight"], P[pre + "in_proj_bias"]
    q = linear(G, Win[:E], bin_[:E])
    k = linear(H, Win[E:2 * E], bin_[E:2 * E])
    v = linear(H, Win[2 * E:], bin_[2 * E:])
    s = (q / np.sqrt(E)) @ k.T
    a = softmax(s, axis=1)
    ctx = a @ v
    out = linear(ctx, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])
    return out, a, (G, H, q, k, v, a, ctx)


def mcat_coattn_bwd(P, cache, dout, grads, pre="co_attention.", dA=None):
    """dA: optional gradient arriving on the returned map (CrossEntropySurvivalAttnRegLoss, loss.py:88-101)."""
    G, H, q, k, v, a, ctx = cache
    E = G.shape[1]
    Win = P[pre + "in_proj_weight"]
    dctx, dWo, dbo = linear_bwd(dout, ctx, P[pre + "out_proj.weight"])
    grads.add(pre + "out_proj.weight", dWo)
    grads.add(pre + "out_proj.bias", dbo)
    da = dctx @ v.T
    if dA is not None:
        da = da + dA
    dv = a.T @ dctx
    ds = softmax_bwd(da, a, axis=1)
    dq = (ds @ k) / np.sqrt(E)
    dk = ds.T @ (q / np.sqrt(E))
    dG, dWq, dbq = linear_bwd(dq, G, Win[:E])
    dHk, dWk, dbk = linear_bwd(dk, H, Win[E:2 * E])
    dHv, dWv, dbv = linear_bwd(dv, H, Win[2 * E:])
    grads.add(pre + "in_proj_weight", np.concatenate([dWq, dWk, dWv], axis=0))
    grads.add(pre + "in_proj_bias", np.concatenate([dbq, dbk, dbv], axis=0))
    return dG, dHk + dHv


# ------------------------------------------------------------------------------------------------ NaCAGaT co-attention
# reference: models/blocks.py:114-206 (pre-gated attention) and :232-253 (contextual attention gate), :92-111
def cag_fwd(P, Q, Qh, pre="co_attention.CAG."):
    f1 = elu(linear(Q, P[pre + "fc1.0.weight"], P[pre + "fc1.0.bias"]))
    f2 = elu(linear(Qh, P[pre + "fc2.0.weight"], P[pre + "fc2.0.bias"]))
    f3 = elu(linear(Qh, P[pre + "fc3.0.weight"], P[pre + "fc3.0.bias"]))
    u = elu(f1 + f2)
    Gg, cG = layernorm(u, P[pre + "G.1.weight"], P[pre + "G.1.bias"])
    w = elu(f3)
    Ee, cE = layernorm(w, P[pre + "E.1.weight"], P[pre + "E.1.bias"])
    m = Gg * Ee
    C = elu(linear(m, P[pre + "fc_c.0.weight"], P[pre + "fc_c.0.bias"]))
    return C, (Q, Qh, f1, f2, f3, u, cG, Gg, w, cE, Ee, m, C)


def cag_bwd(P, cache, dC, grads, pre="co_attention.CAG."):
    Q, Qh, f1, f2, f3, u, cG, Gg, w, cE, Ee, m, C = cache
    dm, dW, db = linear_bwd(elu_bwd(dC, C), m, P[pre + "fc_c.0.weight"])
    grads.add(pre + "fc_c.0.weight", dW)
    grads.add(pre + "fc_c.0.bias", db)
    dGg, dEe = dm * Ee, dm * Gg
    du, dg, dbg = layernorm_bwd(dGg, cG, P[pre + "G.1.weight"])
    grads.add(pre + "G.1.weight", dg)
    grads.add(pre + "G.1.bias", dbg)
    dw, de, dbe = layernorm_bwd(dEe, cE, P[pre + "E.1.weight"])
    grads.add(pre + "E.1.weight", de)
    grads.add(pre + "E.1.bias", dbe)
    df12 = elu_bwd(du, u)
    df3 = elu_bwd(elu_bwd(dw, w), f3)
    dQ, dW1, db1 = linear_bwd(elu_bwd(df12, f1), Q, P[pre + "fc1.0.weight"])
    dQh2, dW2, db2 = linear_bwd(elu_bwd(df12, f2), Qh, P[pre + "fc2.0.weight"])
    dQh3, dW3, db3 = linear_bwd(df3, Qh, P[pre + "fc3.0.weight"])
    for n, (a, b) in {"fc1": (dW1, db1), "fc2": (dW2, db2), "fc3": (dW3, db3)}.items():
        grads.add(pre + n + ".0.weight", a)
        grads.add(pre + n + ".0.bias", b)
    return dQ, dQh2 + dQh3


def nacagat_coattn_fwd(P, G, H, pre="co_attention."):
    E = G.shape[1]
    Win, bin_ = P[pre + "in_proj_weight"], P[pre + "in_proj_bias"]
    q = linear(G, Win[:E], bin_[:E])
    k = linear(H, Win[E:2 * E], bin_[E:2 * E])
    v = linear(H, Win[2 * E:], bin_[2 * E:])
    s = (q / np.sqrt(E)) @ k.T                       # blocks.py:180,184
    tq, tk = np.tanh(q), np.tanh(k)
    Pm = (tq @ tk.T + 1.0) / 2.0                     # blocks.py:185-186
    s2 = s * Pm                                      # blocks.py:187
    a = softmax(s2, axis=1)                          # blocks.py:188
    ctx = a @ v
    out = linear(ctx, P[pre + "out_proj.weight"], P[pre + "out_proj.bias"])
    C, ccag = cag_fwd(P, G, q, pre + "CAG.")          # blocks.py:110  (q, not the attention output: SURVEY F2)
    return out + C, a, (G, H, q, k, v, s, tq, tk, Pm, a, ctx, ccag)


def nacagat_coattn_bwd(P, cache, dout, grads, pre="co_attention.", dA=None):
    G, H, q, k, v, s, tq, tk, Pm, a, ctx, ccag = cache
    E = G.shape[1]
    Win = P[pre + "in_proj_weight"]
    dG_cag, dq_cag = cag_bwd(P, ccag, dout, grads, pre + "CAG.")
    dctx, dWo, dbo = linear_bwd(dout, ctx, P[pre + "out_proj.weight"])
    grads.add(pre + "out_proj.weight", dWo)
    grads.add(pre + "out_proj.bias", dbo)
    da = dctx @ v.T
    if dA is not None:
        da = da + dA
    dv = a.T @ dctx
    ds2 = softmax_bwd(da, a, axis=1)
    ds = ds2 * Pm
    dPm = ds2 * s
    dq = (ds @ k) / np.sqrt(E) + dq_cag
    dk = ds.T @ (q / np.sqrt(E))
    dtq = 0.5 * dPm @ tk
    dtk = 0.5 * dPm.T @ tq
    dq = dq + dtq * (1 - tq * tq)
    dk = dk + dtk * (1 - tk * tk)
    dG, dWq, dbq = linear_bwd(dq, G, Win[:E])
    dHk, dWk, dbk = linear_bwd(dk, H, Win[E:2 * E])
    dHv, dWv, dbv = linear_bwd(dv, H, Win[2 * E:])
    grads.add(pre + "in_proj_weight", np.concatenate([dWq, dWk, dWv], axis=0))
    grads.add(pre + "in_proj_bias", np.concatenate([dbq, dbk, dbv], axis=0))
    return dG + dG_cag, dHk + dHv


# ------------------------------------------------------------------------------------------------ transformer encoder
# reference: models/mcat/mcat.py:51-53,60-62,101-102 -> nn.TransformerEncoderLayer(256, nhead=8, ff=512, relu),
# post-norm, unbatched [L, E] input, no final norm; 2 layers
NHEAD = 8


def encoder_layer_fwd(P, pre, x):
    L, E = x.shape
    hd = E // NHEAD
    qkv = linear(x, P[pre + "self_attn.in_proj_weight"], P[pre + "self_attn.in_proj_bias"])
    q, k, v = [t.reshape(L, NHEAD, hd).transpose(1, 0, 2) for t in np.split(qkv, 3, axis=1)]   # [h, L, hd]
    s = (q @ k.transpose(0, 2, 1)) / np.sqrt(hd)
    a = softmax(s, axis=-1)
    ctx = (a @ v).transpose(1, 0, 2).reshape(L, E)
    sa = linear(ctx, P[pre + "self_attn.out_proj.weight"], P[pre + "self_attn.out_proj.bias"])
    y1, c1 = layernorm(x + sa, P[pre + "norm1.weight"], P[pre + "norm1.bias"])
    f = np.maximum(linear(y1, P[pre + "linear1.weight"], P[pre + "linear1.bias"]), 0.0)
    f2 = linear(f, P[pre + "linear2.weight"], P[pre + "linear2.bias"])
    y2, c2 = layernorm(y1 + f2, P[pre + "norm2.weight"], P[pre + "norm2.bias"])
    return y2, (x, q, k, v, a, ctx, c1, y1, f, c2)


def encoder_layer_bwd(P, pre, cache, dy2, grads):
    x, q, k, v, a, ctx, c1, y1, f, c2 = cache
    L, E = x.shape
    hd = E // NHEAD
    dr2, dg2, db2 = layernorm_bwd(dy2, c2, P[pre + "norm2.weight"])
    grads.add(pre + "norm2.weight", dg2)
    grads.add(pre + "norm2.bias", db2)
    df, dW2, dbl2 = linear_bwd(dr2, f, P[pre + "linear2.weight"])
    grads.add(pre + "linear2.weight", dW2)
    grads.add(pre + "linear2.bias", dbl2)
    df = df * (f > 0)
    dy1, dW1, dbl1 = linear_bwd(df, y1, P[pre + "linear1.weight"])
    grads.add(pre + "linear1.weight", dW1)
    grads.add(pre + "linear1.bias", dbl1)
    dy1 = dy1 + dr2
    dr1, dg1, db1 = layernorm_bwd(dy1, c1, P[pre + "norm1.weight"])
    grads.add(pre + "norm1.weight", dg1)
    grads.add(pre + "norm1.bias", db1)
    dctx, dWo, dbo = linear_bwd(dr1, ctx, P[pre + "self_attn.out_proj.weight"])
    grads.add(pre + "self_attn.out_proj.weight", dWo)
    grads.add(pre + "self_attn.out_proj.bias", dbo)
    dctx_h = dctx.reshape(L, NHEAD, hd).transpose(1, 0, 2)
    da = dctx_h @ v.transpose(0, 2, 1)
    dv = a.transpose(0, 2, 1) @ dctx_h
    ds = softmax_bwd(da, a, axis=-1) / np.sqrt(hd)
    dq = ds @ k
    dk = ds.transpose(0, 2, 1) @ q
    dqkv = np.concatenate([t.transpose(1, 0, 2).reshape(L, E) for t in (dq, dk, dv)], axis=1)
    dx, dWin, dbin = linear_bwd(dqkv, x, P[pre + "self_attn.in_proj_weight"])
    grads.add(pre + "self_attn.in_proj_weight", dWin)
    grads.add(pre + "self_attn.in_proj_bias", dbin)
    return dx + dr1


def encoder_fwd(P, name, x, nlayers=2):
    caches = []
    for l in range(nlayers):
        x, c = encoder_layer_fwd(P, f"{name}.layers.{l}.", x)
        caches.append(c)
    return x, caches


def encoder_bwd(P, name, caches, dy, grads):
    for l in reversed(range(len(caches))):
        dy = encoder_layer_bwd(P, f"{name}.layers.{l}.", caches[l], dy, grads)
    return dy


# ------------------------------------------------------------------------------------------------ gated attention pooling
# reference: models/blocks.py:13-48 (AttentionNetGated) + models/mcat/mcat.py:105-109 (softmax pooling, rho)
def pool_fwd(P, head, rho, x):
    a = np.tanh(linear(x, P[head + ".attention_a.0.weight"], P[head + ".attention_a.0.bias"]))
    b = sigmoid(linear(x, P[head + ".attention_b.0.weight"], P[head + ".attention_b.0.bias"]))
    A = linear(a * b, P[head + ".attention_c.weight"], P[head + ".attention_c.bias"])   # [L,1]
    A_row = A.T                                                                        # [1,L]  (returned, raw)
    w = softmax(A_row, axis=1)
    hp = w @ x                                                                         # [1,E]
    h = np.maximum(linear(hp, P[rho + ".0.weight"], P[rho + ".0.bias"]), 0.0)[0]
    return A_row, h, (x, a, b, w, hp, h)


def pool_bwd(P, head, rho, cache, dh, grads):
    x, a, b, w, hp, h = cache
    dz = (dh * (h > 0))[None, :]
    dhp, dWr, dbr = linear_bwd(dz, hp, P[rho + ".0.weight"])
    grads.add(rho + ".0.weight", dWr)
    grads.add(rho + ".0.bias", dbr)
    dw = dhp @ x.T                       # [1,L]
    dx = w.T @ dhp                       # [L,E]
    dA = softmax_bwd(dw, w, axis=1).T    # [L,1]
    dab, dWc, dbc = linear_bwd(dA, a * b, P[head + ".attention_c.weight"])
    grads.add(head + ".attention_c.weight", dWc)
    grads.add(head + ".attention_c.bias", dbc)
    da = dab * b * (1 - a * a)
    db = dab * a * b * (1 - b)
    dxa, dWa, dba = linear_bwd(da, x, P[head + ".attention_a.0.weight"])
    dxb, dWb, dbb = linear_bwd(db, x, P[head + ".attention_b.0.weight"])
    grads.add(head + ".attention_a.0.weight", dWa)
    grads.add(head + ".attention_a.0.bias", dba)
    grads.add(head + ".attention_b.0.weight", dWb)
    grads.add(head + ".attention_b.0.bias", dbb)
    return dx + dxa + dxb


# ------------------------------------------------------------------------------------------------ fusion
# reference: models/fusion.py:7-19 (concat), :44-113 (bilinear)
def fusion_concat_fwd(P, hp, ho, pre="fusion_layer."):
    c = np.concatenate([hp, ho])
    z1 = np.maximum(linear(c, P[pre + "fusion_layer.0.weight"], P[pre + "fusion_layer.0.bias"]), 0.0)
    z2 = np.maximum(linear(z1, P[pre + "fusion_layer.2.weight"], P[pre + "fusion_layer.2.bias"]), 0.0)
    return z2, (c, z1, z2)


def fusion_concat_bwd(P, cache, dh, grads, pre="fusion_layer."):
    c, z1, z2 = cache
    d2 = dh * (z2 > 0)
    grads.add(pre + "fusion_layer.2.weight", np.outer(d2, z1))
    grads.add(pre + "fusion_layer.2.bias", d2)
    d1 = (d2 @ P[pre + "fusion_layer.2.weight"]) * (z1 > 0)
    grads.add(pre + "fusion_layer.0.weight", np.outer(d1, c))
    grads.add(pre + "fusion_layer.0.bias", d1)
    dc = d1 @ P[pre + "fusion_layer.0.weight"]
    n = c.shape[0] // 2
    return dc[:n], dc[n:]


def _bilinear_gate_fwd(P, pre, idx, xa, xb):
    h = np.maximum(linear(xa, P[f"{pre}linear_h{idx}.0.weight"], P[f"{pre}linear_h{idx}.0.bias"]), 0.0)
    Wz = P[f"{pre}linear_z{idx}.weight"]                                  # [32, 256, 256]
    z = np.einsum("i,kij,j->k", xa, Wz, xb) + P[f"{pre}linear_z{idx}.bias"]
    g = sigmoid(z)
    o = np.maximum(linear(g * h, P[f"{pre}linear_o{idx}.0.weight"], P[f"{pre}linear_o{idx}.0.bias"]), 0.0)
    return o, (xa, xb, h, g, o)


def _bilinear_gate_bwd(P, pre, idx, cache, do, grads):
    xa, xb, h, g, o = cache
    dpre = do * (o > 0)
    grads.add(f"{pre}linear_o{idx}.0.weight", np.outer(dpre, g * h))
    grads.add(f"{pre}linear_o{idx}.0.bias", dpre)
    dgh = dpre @ P[f"{pre}linear_o{idx}.0.weight"]
    dh = dgh * g * (h > 0)
    dz = dgh * h * g * (1 - g)
    Wz = P[f"{pre}linear_z{idx}.weight"]
    grads.add(f"{pre}linear_z{idx}.weight", np.einsum("k,i,j->kij", dz, xa, xb))
    grads.add(f"{pre}linear_z{idx}.bias", dz)
    dxa = np.einsum("k,kij,j->i", dz, Wz, xb)
    dxb = np.einsum("k,kij,i->j", dz, Wz, xa)
    grads.add(f"{pre}linear_h{idx}.0.weight", np.outer(dh, xa))
    grads.add(f"{pre}linear_h{idx}.0.bias", dh)
    dxa = dxa + dh @ P[f"{pre}linear_h{idx}.0.weight"]
    return dxa, dxb


def fusion_bilinear_fwd(P, x1, x2, pre="fusion_layer."):
    o1, c1 = _bilinear_gate_fwd(P, pre, 1, x1, x2)            # fusion.py:88-90
    o2, c2 = _bilinear_gate_fwd(P, pre, 2, x2, x1)            # fusion.py:95-97
    o1e = np.concatenate([o1, [1.0]])
    o2e = np.concatenate([o2, [1.0]])
    kp = np.outer(o1e, o2e).reshape(-1)                       # fusion.py:102-106
    f1 = np.maximum(linear(kp, P[pre + "fc1.0.weight"], P[pre + "fc1.0.bias"]), 0.0)
    cat = np.concatenate([f1, o1e, o2e])                      # fusion.py:110-111
    f2 = np.maximum(linear(cat, P[pre + "fc2.0.weight"], P[pre + "fc2.0.bias"]), 0.0)
    return f2, (c1, c2, o1e, o2e, kp, f1, cat, f2)


def fusion_bilinear_bwd(P, cache, dout, grads, pre="fusion_layer."):
    c1, c2, o1e, o2e, kp, f1, cat, f2 = cache
    d2 = dout * (f2 > 0)
    grads.add(pre + "fc2.0.weight", np.outer(d2, cat))
    grads.add(pre + "fc2.0.bias", d2)
    dcat = d2 @ P[pre + "fc2.0.weight"]
    nf = f1.shape[0]
    n1 = o1e.shape[0]
    df1 = dcat[:nf] * (f1 > 0)
    do1e = dcat[nf:nf + n1].copy()
    do2e = dcat[nf + n1:].copy()
    grads.add(pre + "fc1.0.weight", np.outer(df1, kp))
    grads.add(pre + "fc1.0.bias", df1)
    dkp = (df1 @ P[pre + "fc1.0.weight"]).reshape(n1, n1)
    do1e += dkp @ o2e
    do2e += dkp.T @ o1e
    dx1a, dx2a = _bilinear_gate_bwd(P, pre, 1, c1, do1e[:-1], grads)
    dx2b, dx1b = _bilinear_gate_bwd(P, pre, 2, c2, do2e[:-1], grads)
    return dx1a + dx1b, dx2a + dx2b


# ------------------------------------------------------------------------------------------------ survival head + losses
# reference: models/mcat/mcat.py:126-138
def surv_head_fwd(P, h):
    logits = linear(h, P["classifier.weight"], P["classifier.bias"])[None, :]
    hazards = sigmoid(logits)
    S = np.cumprod(1.0 - hazards, axis=1)
    Y = softmax(logits, axis=1)
    return logits, hazards, S, Y


def surv_head_bwd(P, h, logits, hazards, S, Y, dhaz, dS, dY, grads):
    """gradients w.r.t. hazards [1,K], S [1,K], Y [1,K] -> dh."""
    K = hazards.shape[1]
    dhz = np.array(dhaz, dtype=F64).copy()
    # S_j = prod_{t<=j} (1 - hz_t)  =>  dS_j/dhz_t = -S_j / (1 - hz_t) for t <= j
    for t in range(K):
        dhz[0, t] += -(dS[0, t:] * S[0, t:]).sum() / (1.0 - hazards[0, t])
    dlogits = dhz * hazards * (1 - hazards) + softmax_bwd(np.asarray(dY, F64), Y, axis=1)
    grads.add("classifier.weight", np.outer(dlogits[0], h))
    grads.add("classifier.bias", dlogits[0])
    return dlogits[0] @ P["classifier.weight"]


def fusion_gated_concat_fwd(P, hp, ho, gates, pre="fusion_layer."):
    """reference: models/fusion.py:22-41 (GatedConcatFusion).  gates = [(w [1,256], b [1])] * 2: the reference keeps
    them in a plain Python list, so they are unregistered, untrained and absent from the state_dict (fusion.py:25-27)."""
    items, gcache = [], []
    for (w, b), x in zip(gates, (hp, ho)):
        g = sigmoid(float(np.asarray(w, F64).reshape(-1) @ x + np.asarray(b, F64).reshape(-1)[0]))
        items.append(x * g)
        gcache.append((np.asarray(w, F64).reshape(-1), x, g))
    out, c = fusion_concat_fwd(P, items[0], items[1], pre)
    return out, (c, gcache)


def fusion_gated_concat_bwd(P, cache, dh, grads, pre="fusion_layer."):
    c, gcache = cache
    d_items = fusion_concat_bwd(P, c, dh, grads, pre)
    outs = []
    for (w, x, g), di in zip(gcache, d_items):
        outs.append(di * g + (di @ x) * g * (1.0 - g) * w)        # item = x * sigmoid(w.x + b)
    return outs[0], outs[1]


def sct_loss(Y, label, c, eps=1e-7):
    """reference: models/loss.py:62-85 (SurvivalClassificationTobitLoss).  Returns loss and d/dY."""
    Yv = np.asarray(Y, F64).reshape(-1)
    y = int(np.asarray(label).reshape(-1)[0])
    dY = np.zeros_like(Yv)
    if float(np.asarray(c).reshape(-1)[0]) == 0:
        loss = -np.log(Yv[y] + eps)
        dY[y] = -1.0 / (Yv[y] + eps)
    else:
        cum = Yv[y:].sum() + eps
        loss = -np.log(cum)
        dY[y:] = -1.0 / cum
    return float(loss), dY.reshape(np.asarray(Y).shape)


def attn_norm_reg(A, lambda_reg):
    """reference: models/loss.py:97  lambda_reg * torch.norm(attention, p=2).  Returns the term and d/dA."""
    A = np.asarray(A, F64)
    nrm = np.sqrt((A * A).sum())
    return float(lambda_reg * nrm), (lambda_reg * A / nrm if nrm > 0 else np.zeros_like(A))


def nll_surv_loss(hazards, S, Y, c, alpha=0.15, eps=1e-7):
    """reference: models/loss.py:31-43.  Returns loss, d/dhazards, d/dS (batch of 1)."""
    y = int(np.asarray(Y).reshape(-1)[0])
    c = float(np.asarray(c).reshape(-1)[0])
    hz = np.asarray(hazards, F64).reshape(1, -1)
    S = np.asarray(S, F64).reshape(1, -1)
    Sp = np.concatenate([[1.0], S[0]])
    s_prev, h_y, s_y = Sp[y], hz[0, y], Sp[y + 1]
    unc = -(1 - c) * (np.log(max(s_prev, eps)) + np.log(max(h_y, eps)))
    cen = -c * np.log(max(s_y, eps))
    loss = (1 - alpha) * (cen + unc) + alpha * unc
    dS = np.zeros_like(S)
    dhz = np.zeros_like(hz)
    w_unc = (1 - alpha) + alpha
    if y >= 1 and s_prev > eps:
        dS[0, y - 1] += w_unc * (-(1 - c) / s_prev)
    if h_y > eps:
        dhz[0, y] += w_unc * (-(1 - c) / h_y)
    if s_y > eps:
        dS[0, y] += (1 - alpha) * (-c / s_y)
    return float(loss), dhz, dS


def ces_surv_loss(hazards, S, Y, c, alpha=0.75, eps=1e-7):
    """reference: models/loss.py:5-28 (CrossEntropySurvivalLoss).  Returns loss, d/dhazards, d/dS."""
    y = int(np.asarray(Y).reshape(-1)[0])
    c = float(np.asarray(c).reshape(-1)[0])
    hz = np.asarray(hazards, F64).reshape(1, -1)
    S = np.asarray(S, F64).reshape(1, -1)
    Sp = np.concatenate([[1.0], S[0]])
    s_prev, h_y, s_y = Sp[y], hz[0, y], S[0, y]
    reg = -(1 - c) * (np.log(max(s_prev, eps)) + np.log(max(h_y, eps)))
    s_yc = max(s_y, eps)
    ce = -(c * np.log(s_yc) + (1 - c) * np.log(1 - s_yc))
    loss = (1 - alpha) * ce + alpha * reg
    dS = np.zeros_like(S)
    dhz = np.zeros_like(hz)
    if y >= 1 and s_prev > eps:
        dS[0, y - 1] += alpha * (-(1 - c) / s_prev)
    if h_y > eps:
        dhz[0, y] += alpha * (-(1 - c) / h_y)
    if s_y > eps:
        dS[0, y] += (1 - alpha) * (-(c / s_yc) + (1 - c) / (1 - s_yc))
    return float(loss), dhz, dS


def risk_score(S):
    """reference: models/mcat/main.py:56  risk = -sum(survs)."""
    return -np.asarray(S, F64).sum(axis=1)


# ------------------------------------------------------------------------------------------------ whole models
def _to64(P):
    return {k: np.asarray(v, F64) for k, v in P.items()}


def model_forward_backward(P, wsi, omics, label=None, censor=None, model="mcat", fusion="concat", loss="nll",
                           want_grads=True, gates=None, lambda_reg=0.01):
    """One slide through MCAT (models/mcat/mcat.py:84-142) or NaCAGaT (models/nacagat/nacagat.py:80-138), eval mode.

    Returns a dict with hazards, S, Y, risk, coattn [6,N], path [1,6], omic [1,6], loss and (optionally) grads."""
    P = _to64(P)
    X = np.asarray(wsi, F64)
    if X.ndim == 3:
        X = X[0]
    H = bag_proj_fwd(P, X)
    G, csnn = snn_fwd(P, omics)
    if model == "mcat":
        Hc, A, cco = mcat_coattn_fwd(P, G, H)
    elif model == "nacagat":
        Hc, A, cco = nacagat_coattn_fwd(P, G, H)
    else:
        raise ValueError(model)
    pt, cpt = encoder_fwd(P, "path_transformer", Hc)
    ot, cot = encoder_fwd(P, "omic_transformer", G)
    A_path, h_path, cpp = pool_fwd(P, "path_attention_head", "path_rho", pt)
    A_omic, h_omic, cpo = pool_fwd(P, "omic_attention_head", "omic_rho", ot)
    if fusion == "concat":
        h, cf = fusion_concat_fwd(P, h_path, h_omic)
    elif fusion == "bilinear":
        h, cf = fusion_bilinear_fwd(P, h_path, h_omic)
    elif fusion == "gated_concat":
        h, cf = fusion_gated_concat_fwd(P, h_path, h_omic, gates)
    else:
        raise ValueError(fusion)
    logits, hazards, S, Y = surv_head_fwd(P, h)
    out = dict(hazards=hazards, S=S, Y=Y, risk=risk_score(S), coattn=A, path=A_path, omic=A_omic, logits=logits,
               H_coattn=Hc, G_bag=G)
    if label is None:
        return out
    dY, dA = np.zeros_like(Y), None
    if loss == "nll":
        L, dhz, dS = nll_surv_loss(hazards, S, label, censor)
    elif loss == "ces":
        L, dhz, dS = ces_surv_loss(hazards, S, label, censor)
    elif loss == "sct":
        L, dY = sct_loss(Y, label, censor)
        dhz, dS = np.zeros_like(hazards), np.zeros_like(S)
    elif loss == "cesar":
        L, dhz, dS = ces_surv_loss(hazards, S, label, censor)
        reg, dA = attn_norm_reg(A, lambda_reg)
        L = L + reg
    else:
        raise ValueError(loss)
    out["loss"] = L
    if not want_grads:
        return out
    grads = Grads()
    dh = surv_head_bwd(P, h, logits, hazards, S, Y, dhz, dS, dY, grads)
    if fusion == "concat":
        dhp, dho = fusion_concat_bwd(P, cf, dh, grads)
    elif fusion == "gated_concat":
        dhp, dho = fusion_gated_concat_bwd(P, cf, dh, grads)
    else:
        dhp, dho = fusion_bilinear_bwd(P, cf, dh, grads)
    dpt = pool_bwd(P, "path_attention_head", "path_rho", cpp, dhp, grads)
    dot = pool_bwd(P, "omic_attention_head", "omic_rho", cpo, dho, grads)
    dHc = encoder_bwd(P, "path_transformer", cpt, dpt, grads)
    dG = encoder_bwd(P, "omic_transformer", cot, dot, grads)
    if model == "mcat":
        dG2, dH = mcat_coattn_bwd(P, cco, dHc, grads, dA=dA)
    else:
        dG2, dH = nacagat_coattn_bwd(P, cco, dHc, grads, dA=dA)
    out["_internals"] = dict(H=H, dHc=dHc, coattn_cache=cco, dH=dH)      # for bag-stage level comparisons in tests
    snn_bwd(P, csnn, dG + dG2, grads)
    bag_proj_bwd(P, X, H, dH, grads)
    out["grads"] = dict(grads)
    return out


# ------------------------------------------------------------------------------------------------ GE-NaCAGaT
def ge_forward_backward(P, wsi, label=None, want_grads=True):
    """One slide through GeneExprNarrowContextualAttentionGateTransformer (models/ge_nacagat/ge_nacagat.py:41-72), eval
    mode: H projection, 1-head N x N self-attention over the patches (the averaged map is returned), 2-layer 8-head
    encoder over the N tokens, gated attention pooling over N, 3-class softmax.  The loss is the reference driver's
    nn.CrossEntropyLoss applied to the already soft-maxed Y (models/ge_nacagat/main.py:29,33: a double softmax)."""
    P = _to64(P)
    X = np.asarray(wsi, F64)
    if X.ndim == 3:
        X = X[0]
    N = X.shape[0]
    H = bag_proj_fwd(P, X)
    E = H.shape[1]
    Win, bin_ = P["self_attention.in_proj_weight"], P["self_attention.in_proj_bias"]
    qkv = linear(H, Win, bin_)
    q, k, v = np.split(qkv, 3, axis=1)
    s = (q @ k.T) / np.sqrt(E)
    a = softmax(s, axis=-1)                      # [N,N]; one head, so the head average is the map itself
    ctx = a @ v
    sa = linear(ctx, P["self_attention.out_proj.weight"], P["self_attention.out_proj.bias"])
    pt, cpt = encoder_fwd(P, "path_transformer", sa)
    A_path, h, cpp = pool_fwd(P, "path_attention_head", "path_rho", pt)
    logits = linear(h[None, :], P["classifier.weight"], P["classifier.bias"])[0]
    Y = softmax(logits, axis=-1)
    out = dict(Y=Y, attn=a, path=A_path, logits=logits)
    if label is None:
        return out
    z = softmax(Y, axis=-1)
    out["loss"] = float(-np.log(z[int(label)]))
    if not want_grads:
        return out
    grads = Grads()
    dY = z.copy()
    dY[int(label)] -= 1.0
    dlogits = softmax_bwd(dY, Y, axis=-1)
    dh, dWc, dbc = linear_bwd(dlogits[None, :], h[None, :], P["classifier.weight"])
    grads.add("classifier.weight", dWc)
    grads.add("classifier.bias", dbc)
    dpt = pool_bwd(P, "path_attention_head", "path_rho", cpp, dh[0], grads)
    dsa = encoder_bwd(P, "path_transformer", cpt, dpt, grads)
    dctx, dWo, dbo = linear_bwd(dsa, ctx, P["self_attention.out_proj.weight"])
    grads.add("self_attention.out_proj.weight", dWo)
    grads.add("self_attention.out_proj.bias", dbo)
    da = dctx @ v.T
    dv = a.T @ dctx
    ds = softmax_bwd(da, a, axis=-1) / np.sqrt(E)
    dq = ds @ k
    dk = ds.T @ q
    dqkv = np.concatenate([dq, dk, dv], axis=1)
    dH, dWin, dbin = linear_bwd(dqkv, H, Win)
    grads.add("self_attention.in_proj_weight", dWin)
    grads.add("self_attention.in_proj_bias", dbin)
    bag_proj_bwd(P, X, H, dH, grads)
    out["grads"] = dict(grads)
    return out


# ------------------------------------------------------------------------------------------------ folded bag stage
def folded_bag_stage(W_h, b_h, qk, X):
    """The algebra the CUDA bag kernels implement: returns H, scores [6,N], lse [6], pooled [6,256]."""
    H = np.maximum(np.asarray(X, F64) @ np.asarray(W_h, F64).T + np.asarray(b_h, F64), 0.0)
    s = np.asarray(qk, F64) @ H.T
    m = s.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(s - m).sum(axis=1))
    a = np.exp(s - lse[:, None])
    return H, s, lse, a @ H


def folded_bag_stage_bwd(W_h, b_h, qk, X, dpooled):
    """Autograd of folded_bag_stage w.r.t. the folded queries, H.0.weight and H.0.bias given d(pooled) [6,256]
    (the chain mcat.py:87,97 leaves on the bag side once K/V are folded):
    da = dpooled h, ds = a (da - sum a da), dqk = ds H, dH = a^T dpooled + ds^T qk, dz = dH [z > 0], dW = dz^T X."""
    X = np.asarray(X, F64)
    H, s, lse, pooled = folded_bag_stage(W_h, b_h, qk, X)
    a = np.exp(s - lse[:, None])
    dP = np.asarray(dpooled, F64)
    da = dP @ H.T
    ds = a * (da - (a * da).sum(axis=1, keepdims=True))
    dqk = ds @ H
    dH = a.T @ dP + ds.T @ np.asarray(qk, F64)
    dz = dH * (H > 0)
    return dict(dqk=dqk, dW=dz.T @ X, db=dz.sum(axis=0), pooled=pooled, lse=lse, scores=s)
