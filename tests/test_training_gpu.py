"""GPU checks of the batched drivers: the windowed train epoch against the reference-style per-slide
loop (models/mcat/main.py:30-74), validate / test against per-slide module calls, the pinned window stager and the
HBM-resident bag store."""
import copy
import glob
import os
from importlib import import_module

import numpy as np
import pytest
import torch

from helpers import load_case

pytestmark = pytest.mark.gpu


def _pkg(name):
    return import_module("multimodal-path-omic_b200." + name)


def _net(case):
    synth = _pkg("synth")
    cls = _pkg("mcat").MultimodalCoAttentionTransformer if case["model"] == "mcat" else \
        _pkg("nacagat").NarrowContextualAttentionGateTransformer
    net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=case["fusion"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    return net.cuda()


def _slides(lens, seed0):
    synth = _pkg("synth")
    out = []
    for i, n in enumerate(lens):
        bag, omics, lab, cen = synth.make_slide(seed0 + i, n)
        out.append((float(10 + 3 * i), lab, cen, [torch.from_numpy(o) for o in omics], torch.from_numpy(bag)))
    return out


@pytest.mark.parametrize("model", ["mcat", "nacagat"])
def test_train_epoch_equals_the_reference_style_per_slide_loop(model):
    """5 slides, grad_acc_step 2, two epochs: the windowed epoch (2 + 2 + 1 slides, the odd one carried into the next
    epoch's first optimizer step) must leave the same parameters, epoch loss and risks as the per-slide loop of
    models/mcat/main.py:30-74 run on the module API (eval-mode arithmetic so that both see the same dropout-free path)."""
    tr = _pkg("training")
    loss_mod = _pkg("loss")
    case = load_case(model + "_concat_300")
    slides = _slides([300, 129, 517, 64, 200], 500)
    net_a = _net(case)
    net_b = copy.deepcopy(net_a)
    start = {n: p.detach().clone() for n, p in net_a.named_parameters()}
    # (a) reference-style loop.  SGD, not Adam: the update is linear in the gradient, so the comparison measures the
    # gradients of the two loops (Adam's normalisation turns rounding-level differences of near-zero gradients --
    # atomics order, NaCAGaT's batch-wide fp16 scale of dkg -- into full-size steps)
    net_a.eval()
    opt_a = torch.optim.SGD(net_a.parameters(), lr=0.05)
    ces = loss_mod.CrossEntropySurvivalLoss(alpha=0.75)
    ref_loss, ref_risk = [], []
    count = 0
    for epoch in range(2):
        ep_loss = 0.0
        for months, lab, cen, omics, bag in slides:
            hz, S, Y, _ = net_a(wsi=bag.cuda(), omics=[o.cuda() for o in omics])
            loss = ces(hz, S, torch.tensor([[lab]], device="cuda"), c=torch.tensor([cen], device="cuda"))
            ep_loss += loss.item()
            ref_risk.append(float(-S.sum().item()))
            (loss / 2).backward()
            count += 1
            if count % 2 == 0:
                opt_a.step()
                opt_a.zero_grad()
        ref_loss.append(ep_loss / len(slides))
    # (b) windowed epochs
    opt_b = torch.optim.SGD(net_b.parameters(), lr=0.05)
    runner = tr.EpochRunner(net_b, optimizer=opt_b, loss="ces", grad_acc_step=2)
    got = [runner.train_epoch(slides, train_mode=False) for _ in range(2)]
    assert [g["optimizer_steps"] for g in got] == [2, 3] and runner.pending == 0
    assert np.allclose([g["loss"] for g in got], ref_loss, rtol=2e-5)
    assert np.allclose(np.concatenate([g["risk"] for g in got]), ref_risk, rtol=2e-5)
    assert 0.0 <= got[0]["c_index"] <= 1.0
    umax = max(float((p.detach() - start[n]).norm()) for n, p in net_a.named_parameters())
    for (n, p), (_, q) in zip(net_a.named_parameters(), net_b.named_parameters()):
        da, db = p.detach() - start[n], q.detach() - start[n]
        err = float((da - db).norm()) / max(float(da.norm()), 1e-5 * umax)
        assert err < 5e-3, (n, err)


def test_validate_and_test_match_per_slide_calls(tmp_path):
    tr = _pkg("training")
    loss_mod = _pkg("loss")
    case = load_case("mcat_concat_300")
    slides = _slides([300, 129, 517], 600)
    net = _net(case).eval()
    runner = tr.EpochRunner(net, loss="nll", grad_acc_step=4)
    val = runner.validate(slides, window=2)
    recs = runner.test(slides, output_dir=str(tmp_path), model_name="MCAT", patient="P1", epoch=3, window=2)
    nll = loss_mod.NegativeLogLikelihoodSurvivalLoss()
    losses = []
    with torch.no_grad():
        for i, (months, lab, cen, omics, bag) in enumerate(slides):
            hz, S, Y, att = net(wsi=bag.cuda(), omics=[o.cuda() for o in omics], inference=True)
            losses.append(nll(hz, S, torch.tensor([[lab]], device="cuda"), torch.tensor([cen], device="cuda")).item())
            assert torch.allclose(recs[i]["hazards"], hz, rtol=1e-5) and torch.allclose(recs[i]["coattn"], att["coattn"], rtol=1e-5)
            assert abs(val["risk"][i] + S.sum().item()) < 1e-5
    assert abs(val["loss"] - np.mean(losses)) < 1e-5
    files = sorted(glob.glob(os.path.join(str(tmp_path), "ATTN_MCAT_P1_*_E3_*.pt")))
    assert len(files) == 3
    saved = torch.load(files[1])
    assert saved.shape == (6, 129) and torch.allclose(saved, recs[1]["coattn"].cpu())


def test_window_stager_and_resident_store():
    ing = _pkg("ingest")
    synth = _pkg("synth")
    slides = _slides([300, 129, 517, 64], 700)
    samples = [dict(bag=s[4], omics=s[3], label=s[1], censor=s[2]) for s in slides]
    st = ing.WindowStager("cuda", max_rows=900, max_slides=4, omic_sizes=synth.OMIC_SIZES)
    st.stage(0, samples[:2])
    st.stage(1, samples[2:])
    for k, part in ((0, samples[:2]), (1, samples[2:])):
        bag, om, lab, cen = st.acquire(k)
        want = torch.cat([s["bag"] for s in part]).to(torch.bfloat16)
        assert bag.lengths == tuple(s["bag"].shape[0] for s in part)
        assert torch.equal(bag.x.cpu(), want)
        assert torch.equal(om[3].cpu(), torch.stack([s["omics"][3] for s in part]))
        assert lab.tolist() == [s["label"] for s in part] and cen.tolist() == [s["censor"] for s in part]
        st.release(k)
    with pytest.raises(RuntimeError, match="exceeds"):
        st.stage(2, samples)
    store = ing.ResidentBagStore("cuda", [s["bag"] for s in samples])
    w = store.window([1, 2])
    assert w.x.data_ptr() == store.x[300:].data_ptr() and w.lengths == (129, 517)       # zero-copy view
    g = store.window([3, 0])
    assert torch.equal(g.x.cpu(), torch.cat([samples[3]["bag"], samples[0]["bag"]]).to(torch.bfloat16))
