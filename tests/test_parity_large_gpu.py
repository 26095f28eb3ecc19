"""GPU parity at the BENCHMARKED shapes: the steady-state loops of the persistent bag kernels.

With 148 persistent CTAs a bag of <= 4 096 patches gives every CTA at most one tile, so the fixtures of
test_parity_gpu.py never exercise the shared-memory ring wrap-around, the reuse of the TMEM accumulator stages, a
slide change inside a CTA or the split-K flush over many tiles.  The tests here do: a ragged batch of 9 slides with
153 716 packed rows (1 205 tiles: 8-9 tiles per CTA) through BatchTrainer -- the code path bench.py times -- checked
per slide against the oracle (eval mode), for the default forward kernel and for every selectable variant
(MPO_FWD_CLUSTER=1|4, MPO_FWD_PAIR=1), plus train-mode finite differences of the bag-side parameters (the bag-dropout
mask regenerated in the backward pass).  Reference: models/mcat/mcat.py:84-142, models/nacagat/nacagat.py:80-138.

Tolerances: outputs 1e-3 relative, gradients 1e-2 norm-relative (BASELINE.json north_star)."""
import os
import subprocess
import sys
from importlib import import_module

import numpy as np
import pytest
import torch

from helpers import load_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import mpo_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-3
GRAD_TOL = 1e-2
LENS = [16384, 25088, 20000, 17001, 16385, 30000, 12345, 16384, 129]     # 153 716 rows, 1 205 tiles


def _pkg(name):
    return import_module("multimodal-path-omic_b200." + name)


def _build(case):
    synth = _pkg("synth")
    cls = _pkg("mcat").MultimodalCoAttentionTransformer if case["model"] == "mcat" else \
        _pkg("nacagat").NarrowContextualAttentionGateTransformer
    net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=case["fusion"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    return net.cuda()


def _batch(lens, seed0):
    synth = _pkg("synth")
    bpm = _pkg("bagpass")
    slides = [synth.make_slide(seed0 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    return slides, pb, om, labels, cens


def _nrel(got, ref, floor=0.0):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), floor, 1e-300))


@pytest.mark.parametrize("model", ["mcat", "nacagat"])
def test_large_ragged_batch_matches_oracle(model):
    """9 ragged slides / 153 716 rows in ONE packed step (every CTA owns 8-9 consecutive tiles spanning slide
    boundaries): losses, hazards, attention maps and every parameter gradient against the oracle per slide; for MCAT
    also the bag stage in isolation (scores, lse, pooled; dqk, dW_H, db_H for a given d(pooled))."""
    sp, bpm = _pkg("slidepath"), _pkg("bagpass")
    case = load_case(model + "_concat_16384")
    net = _build(case).eval()
    slides, pb, om, labels, cens = _batch(LENS, 700)
    B = len(LENS)
    assert pb.total_rows >= 150000 and pb.num_tiles // 148 >= 7
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=B)
    tr.zero_grad()
    loss, hz, S = tr.step(pb, om, labels, cens, train=False)
    torch.cuda.synchronize()
    st = tr.last_state
    amap = tr.engine.attention_map(st).cpu().numpy()
    loss, hz, S = loss.cpu().numpy(), hz.cpu().numpy(), S.cpu().numpy()
    got = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in tr.grads.items()}
    qk = st.qk.cpu().numpy()
    pooled, lse = st.bag_ws.pooled.cpu().numpy(), st.bag_ws.lse.cpu().numpy()
    scores = st.bag_ws.scores.cpu().numpy()

    ref_g = None
    worst = {}

    def upd(k, v):
        worst[k] = max(worst.get(k, 0.0), v)

    for b, (bag, omics, lab, cen) in enumerate(slides):
        ref = orc.model_forward_backward(case["state"], bag, omics, lab, cen, model=model, fusion="concat", loss="nll")
        r0, r1 = pb.slide_rows(b)
        upd("hazards", float(np.max(np.abs(hz[b] - ref["hazards"][0]) / np.abs(ref["hazards"][0]))))
        upd("S", float(np.max(np.abs(S[b] - ref["S"][0]) / np.abs(ref["S"][0]))))
        upd("loss", abs(loss[b] - ref["loss"]) / max(1.0, abs(ref["loss"])))
        A, Aref = amap[:, r0:r1].astype(np.float64), ref["coattn"]
        upd("coattn", float(np.max(np.abs(A - Aref) / (np.abs(Aref) + 1e-3 * Aref.max()))))
        if model == "mcat":
            H, s_ref, lse_ref, pooled_ref = orc.folded_bag_stage(case["state"]["H.0.weight"], case["state"]["H.0.bias"],
                                                                 qk[b], bag)
            upd("bag.scores", float(np.max(np.abs(scores[:, r0:r1] - s_ref)) / np.max(np.abs(s_ref))))
            upd("bag.lse", float(np.max(np.abs(lse[b] - lse_ref) / (np.abs(lse_ref) + 1.0))))
            upd("bag.pooled", _nrel(pooled[b], pooled_ref))
        if ref_g is None:
            ref_g = {k: v / B for k, v in ref["grads"].items()}
        else:
            for k, v in ref["grads"].items():
                ref_g[k] = ref_g[k] + v / B
    print(model, {k: "%.2e" % v for k, v in worst.items()})
    for k, v in worst.items():
        assert v < OUT_TOL, (k, v)
    gmax = max(float(np.linalg.norm(v)) for v in ref_g.values())
    errs = sorted(((_nrel(got[k], ref_g[k], floor=1e-5 * gmax), k) for k in ref_g), reverse=True)
    print(model, "grad worst", [(k, "%.2e" % e) for e, k in errs[:5]])
    assert errs[0][0] < GRAD_TOL, errs[:5]

    if model == "mcat":
        # the bag backward in isolation: a given d(pooled) per slide -> dqk per slide, dW_H / db_H summed over slides
        rng = np.random.default_rng(9)
        dP = (rng.standard_normal((B, 6, 256)) * 1e-2).astype(np.float32)
        gw = torch.zeros((256, 1024), device="cuda")
        gb = torch.zeros(256, device="cuda")
        dqk = bpm.bag_backward(pb, st.bag_ws, torch.from_numpy(dP).cuda(), st.qk, gw, gb, drop_p=0.0)
        torch.cuda.synchronize()
        dqk, gw, gb = dqk.cpu().numpy(), gw.cpu().numpy(), gb.cpu().numpy()
        dW_ref, db_ref, e_dqk = 0.0, 0.0, 0.0
        for b, (bag, _, _, _) in enumerate(slides):
            o = orc.folded_bag_stage_bwd(case["state"]["H.0.weight"], case["state"]["H.0.bias"], qk[b], bag, dP[b])
            e_dqk = max(e_dqk, _nrel(dqk[b], o["dqk"]))
            dW_ref, db_ref = dW_ref + o["dW"], db_ref + o["db"]
        e = dict(dqk=e_dqk, dW_H=_nrel(gw, dW_ref), db_H=_nrel(gb, db_ref))
        print("bag backward alone", {k: "%.2e" % v for k, v in e.items()})
        for k, v in e.items():
            assert v < GRAD_TOL, (k, v)


def test_large_batch_graph_replay_matches_eager():
    """the captured step (what bench.py replays) == the eager step at 8-9 tiles per CTA, eval mode."""
    sp = _pkg("slidepath")
    case = load_case("mcat_concat_16384")
    net = _build(case).eval()
    _, pb, om, labels, cens = _batch(LENS, 700)
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=len(LENS))
    tr.zero_grad()
    loss_e, _, _ = tr.step(pb, om, labels, cens, train=False)
    g_e, loss_e = tr.flat_grad.clone(), loss_e.clone()
    g = tr.capture(pb, om, labels, cens, train=False)
    tr.zero_grad()
    loss_g, _, _ = g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(loss_g, loss_e, rtol=1e-6, atol=1e-7)
    assert float((tr.flat_grad - g_e).norm() / g_e.norm()) < 1e-5


@pytest.mark.parametrize("env", [{"MPO_FWD_CLUSTER": "1"}, {"MPO_FWD_CLUSTER": "4"}, {"MPO_FWD_PAIR": "1"},
                                 {"MPO_BWD_REGEN": "1"}],
                         ids=["cluster1", "cluster4", "pair", "bwd_regen"])
def test_forward_kernel_variants_at_large_shapes(env):
    """every selectable build of the forward bag kernel -- and the opt-in backward that regenerates dz inside the
    weight-gradient kernel (MPO_BWD_REGEN=1) -- through the large ragged batch and the 16 384 / 25 088-patch reference
    fixtures (the switches are read once per process, hence the subprocess)."""
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_parity_large_gpu.py"),
                        os.path.join(here, "test_parity_gpu.py"), "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-k",
                        "test_large_ragged_batch_matches_oracle or (test_forward_backward_matches_reference and "
                        "(16384 or 25088))"],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("model", ["mcat", "nacagat"])
def test_train_mode_bag_side_gradients_finite_differences(model):
    """Train mode, 2 slides of 2 500 / 1 300 patches (every dropout site on, bag-dropout mask regenerated inside the
    backward kernels): directional central differences of the step's own loss against the analytic gradients of
    H.0.weight, H.0.bias, co_attention.in_proj_weight / _bias.

    The loss is piecewise smooth (ReLU kinks in the bag projection and the encoders), so a central difference carries an
    O(step) error: it is taken at two steps a factor 4 apart and extrapolated linearly to step 0 (measured on the GPU:
    in_proj_weight 0.345 / 0.388 at steps 0.025 / 0.00625 against 0.401 analytic).  H.0.weight is streamed as a bf16
    copy, so a small dense perturbation is rounded away (fd collapses below step 0.01); its probe moves the 4 096
    entries with the largest gradient by +-2^-9 / +-2^-11 along sign(g), which bf16 represents exactly for |w| < 2^-5."""
    sp = _pkg("slidepath")
    case = load_case(model + "_concat_300")
    net = _build(case).train()
    _, pb, om, labels, cens = _batch([2500, 1300], 900)
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=1)
    SEED = 77123

    def loss_sum():
        tr.zero_grad()
        loss, _, _ = tr.step(pb, om, labels, cens, train=True, seed=SEED)
        return float(loss.double().sum().item())

    base = loss_sum()
    assert abs(loss_sum() - base) < 1e-6                  # same seed: same masks in forward and backward
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    P = dict(net.named_parameters())

    def fd_along(k, d, step):
        old = P[k].data.clone()
        P[k].data.copy_(old + step * d); lp = loss_sum()
        P[k].data.copy_(old - step * d); lm = loss_sum()
        P[k].data.copy_(old)
        return (lp - lm) / (2 * step)

    bad, report = [], []
    for k in ("H.0.weight", "H.0.bias", "co_attention.in_proj_weight", "co_attention.in_proj_bias"):
        g = grads[k]
        if k == "H.0.weight":
            flat = g.reshape(-1)
            top = torch.topk(flat.abs(), 4096).indices
            d = torch.zeros_like(flat)
            d[top] = torch.sign(flat[top])
            d = d.view_as(g)
            steps = (2.0 ** -9, 2.0 ** -11)
        else:
            d = g / g.norm()
            steps = (0.02, 0.005)
        an = float((g.double() * d.double()).sum().item())
        f1, f2 = fd_along(k, d, steps[0]), fd_along(k, d, steps[1])
        fd0 = f2 + (f2 - f1) / 3.0                      # linear extrapolation to step 0 (steps 4 : 1)
        report.append((k, "analytic %.4e" % an, "fd %.4e %.4e -> %.4e" % (f1, f2, fd0)))
        if abs(fd0 - an) > 5e-2 * max(abs(an), abs(fd0)) + 1e-6:
            bad.append((k, an, f1, f2, fd0))
    print(model, report)
    assert not bad, bad
