"""Round-2 diagnostics (GPU): where do the NaCAGaT large-batch gradients and the train-mode directional finite
differences go wrong?  python scripts/gpu_diag_r2.py [nac|fd]"""
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


def build(case):
    cls = pkg("mcat").MultimodalCoAttentionTransformer if case["model"] == "mcat" else pkg("nacagat").NarrowContextualAttentionGateTransformer
    net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=case["fusion"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    return net.cuda()


def batch(lens, seed0):
    slides = [synth.make_slide(seed0 + i, n) for i, n in enumerate(lens)]
    pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).cuda() for s in slides])
    om = [torch.stack([torch.from_numpy(s[1][i]) for s in slides]).cuda() for i in range(6)]
    labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device="cuda")
    cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device="cuda")
    return slides, pb, om, labels, cens


def nrel(a, b, floor=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), floor, 1e-300))


def grads_vs_oracle(case, lens, seed0, tag):
    net = build(case).eval()
    slides, pb, om, labels, cens = batch(lens, seed0)
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=len(lens))
    tr.zero_grad()
    tr.step(pb, om, labels, cens, train=False)
    torch.cuda.synchronize()
    got = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in tr.grads.items()}
    ref = None
    for bag, omics, lab, cen in slides:
        r = orc.model_forward_backward(case["state"], bag, omics, lab, cen, model=case["model"], fusion="concat", loss="nll")
        ref = {k: v / len(lens) for k, v in r["grads"].items()} if ref is None else {k: ref[k] + v / len(lens) for k, v in r["grads"].items()}
    gmax = max(np.linalg.norm(v) for v in ref.values())
    errs = sorted(((nrel(got[k], ref[k], 1e-5 * gmax), k, float(np.linalg.norm(ref[k]))) for k in ref), reverse=True)
    print(tag, lens, [(k, "%.2e" % e, "%.1e" % n) for e, k, n in errs[:6]], flush=True)


def nac():
    case = load_case("nacagat_concat_16384")
    grads_vs_oracle(case, [16384], 712, "A1 one slide 16384 (1 tile/CTA)")
    grads_vs_oracle(case, [30000], 705, "A2 one slide 30000 (2 tiles/CTA)")
    grads_vs_oracle(case, [60000], 731, "A3 one slide 60000 (4 tiles/CTA)")
    grads_vs_oracle(case, [300, 517, 1000, 129, 800, 64, 900, 700, 400], 740, "B  nine small slides")
    grads_vs_oracle(case, [16384, 129], 707, "C1 16384+129")
    grads_vs_oracle(case, [10000, 9000], 750, "C2 10000+9000 (1 tile/CTA, 2 slides)")
    for i, n in enumerate([16384, 25088, 20000, 17001, 16385, 30000, 12345, 16384, 129]):
        grads_vs_oracle(case, [n], 700 + i, "D  slide %d of the failing batch alone" % i)
    case = load_case("mcat_concat_16384")
    grads_vs_oracle(case, [60000], 731, "M  mcat one slide 60000")


def fd():
    for model in ("mcat", "nacagat"):
        for train in (False, True):
            case = load_case(model + "_concat_300")
            net = build(case)
            net.train() if train else net.eval()
            _, pb, om, labels, cens = batch([2500, 1300], 900)
            tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=1)

            def loss_sum():
                tr.zero_grad()
                loss, _, _ = tr.step(pb, om, labels, cens, train=train, seed=77123)
                return float(loss.double().sum().item())
            loss_sum()
            grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
            P = dict(net.named_parameters())
            for k in ("H.0.weight", "H.0.bias", "co_attention.in_proj_weight", "co_attention.in_proj_bias", "G.0.1.0.weight"):
                g = grads[k]
                d = g / g.norm()
                an = float(g.double().norm().item())
                out = []
                for eps in (0.4, 0.1, 0.025, 0.00625):
                    old = P[k].data.clone()
                    P[k].data.copy_(old + eps * d); lp = loss_sum()
                    P[k].data.copy_(old - eps * d); lm = loss_sum()
                    P[k].data.copy_(old)
                    out.append("%.4e" % ((lp - lm) / (2 * eps)))
                print(model, "train" if train else "eval", k, "analytic %.4e" % an, "fd(eps 0.4,0.1,0.025,0.00625)", out, flush=True)
            # block-wise for in_proj_weight: q / k / v rows
            g = grads["co_attention.in_proj_weight"]
            for blk, name in enumerate("qkv"):
                d = torch.zeros_like(g); d[blk * 256:(blk + 1) * 256] = g[blk * 256:(blk + 1) * 256]
                if float(d.norm()) == 0: print(name, "zero grad"); continue
                an = float(d.double().norm().item()); d = d / d.norm()
                out = []
                for eps in (0.1, 0.025, 0.00625):
                    old = P["co_attention.in_proj_weight"].data.clone()
                    P["co_attention.in_proj_weight"].data.copy_(old + eps * d); lp = loss_sum()
                    P["co_attention.in_proj_weight"].data.copy_(old - eps * d); lm = loss_sum()
                    P["co_attention.in_proj_weight"].data.copy_(old)
                    out.append("%.4e" % ((lp - lm) / (2 * eps)))
                print(model, "train" if train else "eval", "in_proj block", name, "analytic %.4e" % an, "fd", out, flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["nac", "fd"]
    if "nac" in which: nac()
    if "fd" in which: fd()
