"""Fused cluster tail vs the per-op tail on the same inputs: every named workspace buffer, outputs and gradients."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, warnings
warnings.filterwarnings("ignore")
from importlib import import_module
pkg = "multimodal-path-omic_b200."
synth = import_module(pkg + "synth"); sp = import_module(pkg + "slidepath"); bpm = import_module(pkg + "bagpass")
lib = import_module(pkg + "_lib")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
train = (sys.argv[3] == "train") if len(sys.argv) > 3 else False
dev = torch.device("cuda", 0)
model_name = sys.argv[4] if len(sys.argv) > 4 else "mcat"
cls = (import_module(pkg + "mcat").MultimodalCoAttentionTransformer if model_name == "mcat"
       else import_module(pkg + "nacagat").NarrowContextualAttentionGateTransformer)
names = ["G", "v", "hc", "cat", "z1", "z2", "logits", "dlogits", "dz1", "dz2", "dhc", "dv", "dG", "dqp"]
for e in ("path0", "path1", "omic0", "omic1"):
    names += ["%s_%s" % (e, s) for s in ("qkv", "probs", "ctx", "y1", "xh1", "rs1", "f", "y2", "xh2", "rs2")]
for p in ("pathpool", "omicpool"):
    names += ["%s_%s" % (p, s) for s in ("a", "b", "w", "hp")]
names += ["cag_" + n for n in ("f1", "f2", "f3", "u", "w", "Gg", "Gxh", "Ee", "Exh", "m", "C")]
names += ["snn_h%d" % i for i in range(6)] + ["snn_dz1_%d" % i for i in range(6)] + ["snn_dz2_%d" % i for i in range(6)]

def run(fused):
    os.environ["MPO_TAIL_FUSED"] = "1" if fused else "0"
    torch.manual_seed(0)
    net = cls(omic_sizes=list(synth.OMIC_SIZES)).to(dev)
    net.train() if train else net.eval()
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=B)
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn((B * N, 1024), generator=g).to(torch.bfloat16).to(dev)
    bag = bpm.PackedBag(x, (N,) * B)
    omics = [torch.randn((B, d), generator=g).to(dev) for d in synth.OMIC_SIZES]
    labels = torch.randint(0, 4, (B,), generator=g).to(dev); censor = torch.randint(0, 2, (B,), generator=g).float().to(dev)
    loss, hz, S = tr.step(bag, omics, labels, censor, train=train, seed=1234)
    torch.cuda.synchronize()
    st = tr.last_state
    out = {"loss": loss.clone(), "hazards": hz.clone(), "S": S.clone(), "att_path": st.att_path.clone(),
           "att_omic": st.att_omic.clone(), "qp": st.qp.clone(), "qk": st.qk.clone(), "dpooled": st.dpooled.clone(),
           "dqk": st.dqk.clone(), "pooled": st.bag_ws.pooled.clone()}
    if st.kc is not None:
        out["kc"] = st.kc.clone()
    if st.dsuma is not None and train:
        out["dsuma"] = st.dsuma.clone()
    for n in names:
        try:
            out["ws." + n] = tr.engine.ws_view(tr.model, st, n).clone()
        except KeyError:
            pass
    for n, gr in tr.grads.items():
        out["grad." + n] = gr.clone()
    return out

a = run(False)
b = run(True)
bad = 0
for k in a:
    x, y = a[k].double().flatten(), b[k].double().flatten()
    den = x.abs().max().item() + 1e-30
    err = (x - y).abs().max().item() / den
    nan = bool(torch.isnan(y).any())
    flag = "" if (err < 2e-4 and not nan) else "   <<<<<<"
    if flag:
        bad += 1
    if flag or os.environ.get("CMP_ALL"):
        print("%-60s max|x|=%.3e  rel err %.3e%s%s" % (k, den, err, " NaN" if nan else "", flag))
        if os.environ.get("CMP_VALS") and flag:
            print("    ref  ", [float("%.3e" % v) for v in x[:6].tolist()], " nz=%d/%d" % (int((x != 0).sum()), x.numel()))
            print("    fused", [float("%.3e" % v) for v in y[:6].tolist()], " nz=%d/%d" % (int((y != 0).sum()), y.numel()))
print("compared %d tensors, %d differ (B=%d N=%d train=%s)" % (len(a), bad, B, N, train))
