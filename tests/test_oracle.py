"""The oracle (oracle/mpo_oracle.py) against the reference: golden fixtures generated from the unmodified
reference modules (tests/golden/make_golden.py) and the reference's own known-answer loss test."""
import os
import sys

import numpy as np
import pytest

from helpers import GOLDEN_DIR, alt_cases, digest_errors, ge_cases, golden_cases, load_case, load_ge_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import mpo_oracle as orc  # noqa: E402

OUT_TOL = 2e-4     # reference runs in fp32, the oracle in fp64
GRAD_TOL = 1e-3    # norm-relative, with the noise floor of helpers.digest_errors


@pytest.mark.parametrize("name", golden_cases() + golden_cases(unrounded=True))
def test_oracle_matches_reference_fixture(name):
    c = load_case(name)
    out = orc.model_forward_backward(c["state"], c["bag"], c["omics"], c["label"], c["censor"], model=c["model"],
                                     fusion=c["fusion"], loss="nll")
    g = c["gold"]
    for k in ("hazards", "S", "Y", "risk", "coattn", "path", "omic"):
        ref = g[k]
        err = np.max(np.abs(out[k] - ref) / (np.abs(ref) + 1e-9))
        assert err < OUT_TOL, (k, err)
    assert abs(out["loss"] - float(g["loss_nll"])) < 1e-5
    ces = orc.model_forward_backward(c["state"], c["bag"], c["omics"], c["label"], c["censor"], model=c["model"],
                                     fusion=c["fusion"], loss="ces", want_grads=False)
    assert abs(ces["loss"] - float(g["loss_ces"])) < 1e-5
    worst, details = digest_errors(c, out["grads"])
    # the un-sharpened (sharpen 1.0) fixtures have a flat co-attention: the six path tokens nearly coincide and the
    # reference's fp32 gradients of the path pooling head are ~1e-7 in norm, i.e. at its own rounding noise
    tol = 3 * GRAD_TOL if "unrounded" in name and "sharp" not in name else GRAD_TOL
    assert worst < tol, sorted(details, key=lambda d: -d[2])[:3]
    assert set(out["grads"].keys()) == set(c["param_names"])


def alt_gates(c):
    g = c["gold"]
    return [(g["gate%d_w" % i], g["gate%d_b" % i]) for i in range(2)] if "gate0_w" in g else None


@pytest.mark.parametrize("name", alt_cases())
def test_oracle_matches_reference_alt_fixture(name):
    """the other loss / fusion branches of the reference drivers (sct, cesar, gated_concat): loss value and every
    parameter gradient of that loss against fixtures generated from the unmodified reference."""
    c = load_case(name)
    g = c["gold"]
    out = orc.model_forward_backward(c["state"], c["bag"], c["omics"], c["label"], c["censor"], model=c["model"],
                                     fusion=c["fusion"], loss=str(g["loss_kind"]), gates=alt_gates(c),
                                     lambda_reg=float(g["lambda_reg"]))
    for k in ("hazards", "S", "Y", "coattn"):
        assert np.max(np.abs(out[k] - g[k]) / (np.abs(g[k]) + 1e-9)) < OUT_TOL, k
    assert abs(out["loss"] - float(g["loss"])) < 1e-5
    worst, details = digest_errors(c, out["grads"])
    assert worst < GRAD_TOL, sorted(details, key=lambda d: -d[2])[:3]


@pytest.mark.parametrize("name", ge_cases())
def test_oracle_matches_reference_ge_fixture(name):
    """GE-NaCAGaT (models/ge_nacagat/ge_nacagat.py) forward + the driver's double-softmax cross-entropy + gradients."""
    from importlib import import_module
    synth = import_module("multimodal-path-omic_b200.synth")
    c = load_ge_case(name)
    out = orc.ge_forward_backward(c["state"], c["bag"], c["label"])
    g = c["gold"]
    assert np.max(np.abs(out["Y"] - g["Y"]) / np.abs(g["Y"])) < OUT_TOL
    assert np.max(np.abs(out["path"] - g["path"])) / np.max(np.abs(g["path"])) < OUT_TOL
    A = out["attn"]
    assert A.shape == (c["n"], c["n"])
    assert np.max(np.abs(A[:64, :64] - g["attn_corner"]) / (np.abs(g["attn_corner"]) + 1e-3 * g["attn_corner"].max())) < 1e-3
    assert np.max(np.abs(A.max(axis=1) - g["attn_rowmax"]) / g["attn_rowmax"]) < 1e-3
    dref = g["attn_digest"]
    dgot = synth.grad_digest("attn", A)
    assert abs(dgot[0] - dref[0]) / dref[0] < 1e-4
    assert abs(out["loss"] - float(g["loss"])) < 1e-5
    worst, details = digest_errors(c, out["grads"])
    # the reference's fp32 softmax-over-N reductions leave ~2e-3 of noise on its smallest gradients (norm 6e-6)
    assert worst < 5 * GRAD_TOL, sorted(details, key=lambda d: -d[2])[:3]
    assert set(out["grads"].keys()) == set(c["param_names"])


def test_reference_known_answer_ces_loss():
    # models/loss.py:108-121: exact float equality in the reference's own test
    hz = np.array([[0.51, 0.52, 0.49, 0.48]], np.float32)
    S = np.array([[0.5, 0.4, 0.2, 0.1]], np.float32)
    l0, _, _ = orc.ces_surv_loss(hz, S, [0], [0.0])
    l1, _, _ = orc.ces_surv_loss(hz, S, [0], [1.0])
    assert np.float32(l0) == np.float32(0.6782951951026917)
    assert np.float32(l1) == np.float32(0.1732867956161499)


def test_loss_table_from_reference():
    z = np.load(os.path.join(GOLDEN_DIR, "loss_known_answers.npz"))
    for y, c, nll, ces in z["table"]:
        ln, _, _ = orc.nll_surv_loss(z["hazards"], z["S"], [int(y)], [c])
        lc, _, _ = orc.ces_surv_loss(z["hazards"], z["S"], [int(y)], [c])
        assert abs(ln - nll) < 1e-6 and abs(lc - ces) < 1e-6


def test_loss_gradients_finite_difference():
    rng = np.random.default_rng(0)
    for fn in (orc.nll_surv_loss, orc.ces_surv_loss):
        for y in range(4):
            for c in (0.0, 1.0):
                hz = rng.uniform(0.2, 0.8, size=(1, 4))
                S = np.cumprod(1 - hz, axis=1)
                _, dhz, dS = fn(hz, S, [y], [c])
                for arr, d in ((hz, dhz), (S, dS)):
                    for j in range(4):
                        a = arr.copy(); a[0, j] += 1e-6
                        b = arr.copy(); b[0, j] -= 1e-6
                        lp = fn(a if arr is hz else hz, a if arr is S else S, [y], [c])[0]
                        lm = fn(b if arr is hz else hz, b if arr is S else S, [y], [c])[0]
                        assert abs((lp - lm) / 2e-6 - d[0, j]) < 1e-5


def test_folded_bag_stage_equals_unfolded_attention():
    # folding W_k into the query and W_v behind the pooled vector is exact
    c = load_case("mcat_concat_sharp_517")
    P = {k: np.asarray(v, np.float64) for k, v in c["state"].items()}
    X = np.asarray(c["bag"], np.float64)
    H = orc.bag_proj_fwd(P, X)
    G, _ = orc.snn_fwd(P, c["omics"])
    out, A, _ = orc.mcat_coattn_fwd(P, G, H)
    E = 256
    Win, b_in = P["co_attention.in_proj_weight"], P["co_attention.in_proj_bias"]
    q = G @ Win[:E].T + b_in[:E]
    qk = (q @ Win[E:2 * E]) / 16.0
    _, s, lse, pooled = orc.folded_bag_stage(P["H.0.weight"], P["H.0.bias"], qk, X)
    A2 = np.exp(s - lse[:, None])
    out2 = (pooled @ Win[2 * E:].T + b_in[2 * E:]) @ P["co_attention.out_proj.weight"].T + P["co_attention.out_proj.bias"]
    assert np.max(np.abs(A - A2)) < 1e-12
    assert np.max(np.abs(out - out2)) < 1e-10


def test_folded_bag_stage_backward_finite_difference():
    """oracle.folded_bag_stage_bwd (the checker of the large-shape GPU tests) against central differences of
    sum(pooled * dpooled) in fp64."""
    rng = np.random.default_rng(5)
    n, din, d = 37, 24, 16
    X = rng.standard_normal((n, din))
    W = rng.standard_normal((d, din)) / np.sqrt(din)
    b = rng.standard_normal(d) * 0.1
    qk = rng.standard_normal((6, d)) * 0.7
    dP = rng.standard_normal((6, d))
    out = orc.folded_bag_stage_bwd(W, b, qk, X, dP)

    def f(W_, b_, qk_):
        return float((orc.folded_bag_stage(W_, b_, qk_, X)[3] * dP).sum())
    eps = 1e-6
    for name, arr, grad in (("W", W, out["dW"]), ("b", b, out["db"]), ("qk", qk, out["dqk"])):
        for _ in range(6):
            idx = tuple(int(rng.integers(0, s)) for s in arr.shape)
            old = arr[idx]
            arr[idx] = old + eps; fp = f(W, b, qk)
            arr[idx] = old - eps; fm = f(W, b, qk)
            arr[idx] = old
            fd = (fp - fm) / (2 * eps)
            assert abs(fd - grad[idx]) < 1e-6 * max(1.0, abs(fd)), (name, idx, fd, grad[idx])
