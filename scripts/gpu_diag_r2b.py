"""NaCAGaT: GPU intermediates against the oracle's for single slides (which stage loses the precision?)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import warnings; warnings.filterwarnings("ignore")
from importlib import import_module
import mpo_oracle as orc
from helpers import load_case
pkg = lambda n: import_module("multimodal-path-omic_b200." + n)
synth, sp, bpm = pkg("synth"), pkg("slidepath"), pkg("bagpass")
from gpu_diag_r2 import build, batch, nrel


def one(case, n, seed, fused="1"):
    os.environ["MPO_TAIL_FUSED"] = fused
    net = build(case).eval()
    slides, pb, om, labels, cens = batch([n], seed)
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=1)
    tr.zero_grad()
    tr.step(pb, om, labels, cens, train=False)
    torch.cuda.synchronize()
    st = tr.last_state
    bag, omics, lab, cen = slides[0]
    r = orc.model_forward_backward(case["state"], bag, omics, lab, cen, model="nacagat", fusion="concat", loss="nll")
    I = r["_internals"]
    G, H, q, k, v, s, tq, tk, Pm, a, ctx, ccag = I["coattn_cache"]
    P = {kk: np.asarray(vv, np.float64) for kk, vv in case["state"].items()}
    Win = P["co_attention.in_proj_weight"]; E = 256
    Wk, Wv = Win[E:2 * E], Win[2 * E:]
    pooled_ref = a @ H
    dHc = I["dHc"]
    dctx = dHc @ P["co_attention.out_proj.weight"]
    dpooled_ref = dctx @ Wv
    da = dctx @ v.T
    ds2 = a * (da - (a * da).sum(axis=1, keepdims=True))
    ds = ds2 * Pm
    dPm = ds2 * s
    dqk_ref = ds @ H
    dtq_ref = 0.5 * dPm @ tk
    dkc_ref = ds.sum(axis=1)
    g = lambda t: t.detach().cpu().numpy().astype(np.float64)
    print("---- N=%d seed=%d fused=%s  a.max per query %s" % (n, seed, fused, np.round(a.max(axis=1), 3)))
    print("fwd: pooled nrel %.2e  maxabs/max %.2e | lse %.2e | hazards %.2e | Hc(out) via qp? " % (
        nrel(g(st.bag_ws.pooled)[0], pooled_ref), np.max(np.abs(g(st.bag_ws.pooled)[0] - pooled_ref)) / np.max(np.abs(pooled_ref)),
        np.max(np.abs(g(st.bag_ws.lse)[0] - (np.log(np.exp(s * Pm - (s * Pm).max(axis=1, keepdims=True)).sum(axis=1)) + (s * Pm).max(axis=1)))),
        np.max(np.abs(g(st.hazards)[0] - r["hazards"][0]) / r["hazards"][0])))
    print("fwd per query pooled nrel", ["%.1e" % nrel(g(st.bag_ws.pooled)[0][i], pooled_ref[i]) for i in range(6)])
    print("bwd: dpooled %.2e  dqk %.2e  dtq %.2e  dkc %.2e" % (
        nrel(g(st.dpooled)[0], dpooled_ref), nrel(g(st.dqk)[0], dqk_ref), nrel(g(st.dtq)[0], dtq_ref), nrel(g(st.dkc)[0], dkc_ref)))
    print("bwd per query dpooled", ["%.1e" % nrel(g(st.dpooled)[0][i], dpooled_ref[i]) for i in range(6)],
          "dqk", ["%.1e" % nrel(g(st.dqk)[0][i], dqk_ref[i]) for i in range(6)],
          "dtq", ["%.1e" % nrel(g(st.dtq)[0][i], dtq_ref[i]) for i in range(6)])
    got = {kk: g(vv) for kk, vv in tr.grads.items()}
    gmax = max(np.linalg.norm(vv) for vv in r["grads"].values())
    errs = sorted(((nrel(got[kk], r["grads"][kk], 1e-5 * gmax), kk) for kk in r["grads"]), reverse=True)
    print("grads:", [(kk, "%.1e" % e) for e, kk in errs[:14]], flush=True)


case = load_case("nacagat_concat_16384")
one(case, 16384, 712)
one(case, 16384, 700)
one(case, 16384, 700, fused="0")
one(case, 30000, 705)
one(case, 60000, 731)
