"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (needs /root/reference).

Run from the repository root in the build container:   python tests/golden/make_golden.py
The reference cannot travel to the GPU box, these small fixtures do.  Weights and inputs are not stored: both
come from multimodal-path-omic_b200/synth.py (numpy PCG64), which tests re-run to rebuild the identical case.
Stored per case: hazards, S, Y, risk, attention maps, NLL and CES losses, and a digest (norm, random projection,
16 samples) of every parameter gradient of the NLL loss, from the reference's own autograd in eval() mode.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

CASES = [
    # name,             model,     fusion,     N,    seed, sharpen
    ("mcat_concat_300", "mcat", "concat", 300, 1, 1.0),
    ("mcat_concat_sharp_517", "mcat", "concat", 517, 2, 8.0),
    ("mcat_concat_4096", "mcat", "concat", 4096, 3, 4.0),
    ("mcat_concat_2", "mcat", "concat", 2, 4, 1.0),
    ("mcat_concat_128", "mcat", "concat", 128, 5, 8.0),
    ("mcat_bilinear_200", "mcat", "bilinear", 200, 6, 4.0),
    ("nacagat_concat_300", "nacagat", "concat", 300, 7, 1.0),
    ("nacagat_concat_sharp_517", "nacagat", "concat", 517, 8, 8.0),
    ("nacagat_bilinear_200", "nacagat", "bilinear", 200, 9, 4.0),
    ("nacagat_concat_4096", "nacagat", "concat", 4096, 10, 4.0),
    # the benchmarked shape (BASELINE configs 2/3: 16 384 patches = 128 tiles) and one 25 088-patch shard of config 5
    ("mcat_concat_16384", "mcat", "concat", 16384, 11, 4.0),
    ("nacagat_concat_16384", "nacagat", "concat", 16384, 12, 4.0),
    ("mcat_concat_25088", "mcat", "concat", 25088, 13, 6.0),
    # SURVEY H2 option (a), second half: the reference run on weights that are NOT bf16-representable (the CUDA path
    # rounds H.0.weight / W_k itself): the delta is reported by tests/test_parity_gpu.py::test_unrounded_weights_delta
    # (sharpen 1.0 = the statistics of the reference's own initialisation; the `sharp` one scales the in-projection by 4)
    ("mcat_concat_unrounded_4096", "mcat", "concat", 4096, 14, 1.0),
    ("nacagat_concat_unrounded_4096", "nacagat", "concat", 4096, 15, 1.0),
    ("nacagat_concat_unrounded_sharp_4096", "nacagat", "concat", 4096, 15, 4.0),
]


# other loss branches of the reference drivers (models/nacagat/main.py:41-50): gradient digests of `loss_kind`
# name,  model, fusion, N, seed, sharpen, loss kind, lambda_reg
ALT_CASES = [
    ("alt_mcat_concat_sct_300", "mcat", "concat", 300, 17, 2.0, "sct", 0.0),          # label 1, censored
    ("alt_mcat_concat_sct_517", "mcat", "concat", 517, 18, 2.0, "sct", 0.0),          # label 2, uncensored
    ("alt_nacagat_concat_cesar_300", "nacagat", "concat", 300, 19, 2.0, "cesar", 5.0),
    ("alt_mcat_concat_cesar_517", "mcat", "concat", 517, 20, 4.0, "cesar", 5.0),
    ("alt_mcat_gated_concat_300", "mcat", "gated_concat", 300, 24, 2.0, "nll", 0.0),
    ("alt_nacagat_gated_concat_517", "nacagat", "gated_concat", 517, 25, 2.0, "ces", 0.0),
]


GE_CASES = [
    # name,        N,   seed, sharpen
    ("ge_300", 300, 21, 1.0),
    ("ge_sharp_517", 517, 22, 4.0),
    ("ge_1000", 1000, 23, 2.0),
    ("ge_4096", 4096, 26, 2.0),       # 32 x 32 output tiles per attention product: the persistent tensor-core GEMM loops
]


def import_reference():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))      # models/utils.py:1 imports it, unused on this path
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from models.mcat.mcat import MultimodalCoAttentionTransformer
    from models.nacagat.nacagat import NarrowContextualAttentionGateTransformer
    from models.loss import NegativeLogLikelihoodSurvivalLoss, CrossEntropySurvivalLoss
    return (MultimodalCoAttentionTransformer, NarrowContextualAttentionGateTransformer,
            NegativeLogLikelihoodSurvivalLoss, CrossEntropySurvivalLoss)


def main():
    from importlib import import_module
    synth = import_module("multimodal-path-omic_b200.synth")
    MCAT, NACAGAT, NLL, CES = import_reference()
    torch.set_num_threads(8)
    outdir = os.path.dirname(os.path.abspath(__file__))
    only = set(sys.argv[1:])          # optional: generate only the named cases (existing fixtures stay untouched)
    for name, model, fusion, n, seed, sharpen in CASES:
        if only and name not in only:
            continue
        cls = MCAT if model == "mcat" else NACAGAT
        torch.manual_seed(seed)
        net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=fusion)
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        state = synth.make_state(shapes, seed, model=model, sharpen=sharpen, round_bag_weights="unrounded" not in name)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
        net.eval()
        bag, omics, label, censor = synth.make_slide(seed, n)
        wsi = torch.from_numpy(bag)
        om = [torch.from_numpy(o) for o in omics]
        Y_t = torch.tensor([[label]], dtype=torch.int64)
        c_t = torch.tensor([censor])
        if model == "mcat":
            hazards, S, Y, att = net(wsi=wsi, omics=om, inference=True)
        else:
            hazards, S, Y, att = net(wsi=wsi, omics=om)
        loss_nll = NLL()(hazards, S, Y_t, c_t)
        loss_ces = CES()(hazards, S, Y_t, c=c_t)
        net.zero_grad()
        loss_nll.backward()
        rec = dict(
            hazards=hazards.detach().numpy(), S=S.detach().numpy(), Y=Y.detach().numpy(),
            risk=(-torch.sum(S, dim=1)).detach().numpy(),
            coattn=att["coattn"].detach().numpy(), path=att["path"].detach().numpy(),
            omic=att["omic"].detach().numpy(),
            loss_nll=np.float64(loss_nll.item()), loss_ces=np.float64(loss_ces.item()),
            meta=np.array([n, seed, label, censor, sharpen], dtype=np.float64),
        )
        names = []
        for k, p_ in net.named_parameters():
            g = p_.grad.detach().numpy() if p_.grad is not None else np.zeros(tuple(p_.shape), np.float32)
            rec["gd/" + k] = synth.grad_digest(k, g)
            names.append(k)
        rec["param_names"] = np.array(names)
        rec["param_shapes"] = np.array([str(shapes[k]) for k in names])
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **rec)
        print(f"{name}: hazards={rec['hazards'].round(5).tolist()} loss_nll={rec['loss_nll']:.6f} "
              f"coattn max={rec['coattn'].max():.3e} min={rec['coattn'].min():.3e}")

    from models.loss import SurvivalClassificationTobitLoss, CrossEntropySurvivalAttnRegLoss
    for name, model, fusion, n, seed, sharpen, kind, lam in ALT_CASES:
        if only and name not in only:
            continue
        cls = MCAT if model == "mcat" else NACAGAT
        torch.manual_seed(seed)
        net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=fusion)
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        state = synth.make_state(shapes, seed, model=model, sharpen=sharpen)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
        net.eval()
        rec = {}
        if fusion == "gated_concat":
            # the reference keeps the gates in a plain Python list (fusion.py:25-27): unregistered, absent from the
            # state_dict; their torch-default initial values are stored with the fixture
            for gi, gate in enumerate(net.fusion_layer.gates):
                rec["gate%d_w" % gi] = gate[0].weight.detach().numpy().copy()
                rec["gate%d_b" % gi] = gate[0].bias.detach().numpy().copy()
        bag, omics, label, censor = synth.make_slide(seed, n)
        wsi = torch.from_numpy(bag)
        om = [torch.from_numpy(o) for o in omics]
        Y_t = torch.tensor([[label]], dtype=torch.int64)
        c_t = torch.tensor([censor])
        hazards, S, Y, att = net(wsi=wsi, omics=om, inference=True) if model == "mcat" else net(wsi=wsi, omics=om)
        if kind == "sct":
            loss = SurvivalClassificationTobitLoss()(Y, Y_t.reshape(1), c=c_t)
        elif kind == "cesar":
            loss, _ = CrossEntropySurvivalAttnRegLoss(lambda_reg=lam)(hazards, S, Y_t, c=c_t, attention=att["coattn"])
        elif kind == "nll":
            loss = NLL()(hazards, S, Y_t, c_t)
        else:
            loss = CES()(hazards, S, Y_t, c=c_t)
        net.zero_grad()
        loss.backward()
        rec.update(hazards=hazards.detach().numpy(), S=S.detach().numpy(), Y=Y.detach().numpy(),
                   coattn=att["coattn"].detach().numpy(), loss=np.float64(loss.item()),
                   meta=np.array([n, seed, label, censor, sharpen], dtype=np.float64), lambda_reg=np.float64(lam),
                   loss_kind=np.array(kind))
        names = []
        for k, p_ in net.named_parameters():
            g = p_.grad.detach().numpy() if p_.grad is not None else np.zeros(tuple(p_.shape), np.float32)
            rec["gd/" + k] = synth.grad_digest(k, g)
            names.append(k)
        rec["param_names"] = np.array(names)
        rec["param_shapes"] = np.array([str(shapes[k]) for k in names])
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **rec)
        print(f"{name}: loss[{kind}]={rec['loss']:.6f} hazards={rec['hazards'].round(5).tolist()}")

    # GE-NaCAGaT (models/ge_nacagat/ge_nacagat.py): Y, the N x N self-attention map (digest + corner block), the
    # pooling logits, the driver's CrossEntropyLoss on Y (models/ge_nacagat/main.py:33) and its gradient digests
    from models.ge_nacagat.ge_nacagat import GeneExprNarrowContextualAttentionGateTransformer as GE
    for name, n, seed, sharpen in GE_CASES:
        if only and name not in only:
            continue
        torch.manual_seed(seed)
        net = GE()
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        state = synth.make_state(shapes, seed, model="ge", sharpen=sharpen)
        net.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()})
        net.eval()
        bag, _, label, _ = synth.make_slide(seed, n)
        label = label % 3
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            Y, att = net(wsi=torch.from_numpy(bag))
        loss = torch.nn.CrossEntropyLoss()(Y.unsqueeze(0), torch.tensor([label]))
        net.zero_grad()
        loss.backward()
        A = att["attn"].detach().numpy()
        rec = dict(Y=Y.detach().numpy(), path=att["path"].detach().numpy(), attn_corner=A[:64, :64].copy(),
                   attn_digest=synth.grad_digest("attn", A), attn_rowmax=A.max(axis=1), loss=np.float64(loss.item()),
                   meta=np.array([n, seed, label, 0.0, sharpen], dtype=np.float64))
        names = []
        for k, p_ in net.named_parameters():
            g = p_.grad.detach().numpy() if p_.grad is not None else np.zeros(tuple(p_.shape), np.float32)
            rec["gd/" + k] = synth.grad_digest(k, g)
            names.append(k)
        rec["param_names"] = np.array(names)
        rec["param_shapes"] = np.array([str(shapes[k]) for k in names])
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **rec)
        print(f"{name}: Y={rec['Y'].round(5).tolist()} loss={rec['loss']:.6f} attn max={A.max():.3e}")

    if only and "loss_known_answers" not in only:
        return
    # loss known answers: the reference's own test vectors (models/loss.py:108-121) plus the SURVEY 8c probes
    hz = torch.tensor([0.51, 0.52, 0.49, 0.48]).reshape(1, 4)
    S = torch.tensor([0.5, 0.4, 0.2, 0.1]).reshape(1, 4)
    rows = []
    for y in range(4):
        for c in (0.0, 1.0):
            if y + 1 > 4:
                continue
            ln = NLL()(hz, S, torch.tensor([y]), torch.tensor([c])).item() if y < 4 else float("nan")
            lc = CES()(hz, S, torch.tensor([y]), torch.tensor([c])).item()
            rows.append([y, c, ln, lc])
    np.savez_compressed(os.path.join(outdir, "loss_known_answers.npz"), hazards=hz.numpy(), S=S.numpy(),
                        table=np.array(rows, dtype=np.float64))
    print("loss table:\n", np.array(rows))


if __name__ == "__main__":
    main()
