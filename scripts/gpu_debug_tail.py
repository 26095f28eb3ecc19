"""Stage-by-stage comparison of the CUDA slide path with the oracle (debug aid, run under gpurun)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from importlib import import_module
from helpers import load_case, digest_errors
import mpo_oracle as orc
from test_parity_gpu import build_model
sp = import_module("multimodal-path-omic_b200.slidepath"); bpm = import_module("multimodal-path-omic_b200.bagpass")

name = sys.argv[1] if len(sys.argv) > 1 else "mcat_concat_300"
case = load_case(name)
net = build_model(case).eval()
eng = net._engine
model = eng.binding.build(grads=None)
bag = bpm.PackedBag.from_slides([torch.from_numpy(case["bag"]).cuda()])
st = eng.forward(model, bag, [torch.from_numpy(o).cuda().reshape(1, -1) for o in case["omics"]], train=False)
torch.cuda.synchronize()
P = {k: np.asarray(v, np.float64) for k, v in case["state"].items()}
X = np.asarray(case["bag"], np.float64)
H = orc.bag_proj_fwd(P, X); G, _ = orc.snn_fwd(P, case["omics"])
if case["model"] == "mcat": Hc, A, _ = orc.mcat_coattn_fwd(P, G, H)
else: Hc, A, _ = orc.nacagat_coattn_fwd(P, G, H)
pt, _ = orc.encoder_fwd(P, "path_transformer", Hc); ot, _ = orc.encoder_fwd(P, "omic_transformer", G)
def cmp(tag, got, ref):
    got = got.detach().cpu().numpy().astype(np.float64).reshape(ref.shape)
    print(f"{tag:12s} max|ref|={np.abs(ref).max():.3e} maxerr={np.abs(got-ref).max():.3e} rel={np.linalg.norm(got-ref)/(np.linalg.norm(ref)+1e-30):.3e}")
cmp("G", eng.ws_view(model, st, "G"), G)
E = 256; Win, b_in = P["co_attention.in_proj_weight"], P["co_attention.in_proj_bias"]
q = G @ Win[:E].T + b_in[:E]; cmp("qp", st.qp, q); cmp("qk", st.qk, (q @ Win[E:2*E]) / 16)
cmp("hc", eng.ws_view(model, st, "hc"), Hc)
cmp("path0_y2", eng.ws_view(model, st, "path0_y2"), orc.encoder_layer_fwd(P, "path_transformer.layers.0.", Hc)[0])
cmp("path1_y2", eng.ws_view(model, st, "path1_y2"), pt)
cmp("omic1_y2", eng.ws_view(model, st, "omic1_y2"), ot)
out = orc.model_forward_backward(case["state"], case["bag"], case["omics"], case["label"], case["censor"], model=case["model"], fusion=case["fusion"])
cmp("hazards", st.hazards, out["hazards"]); cmp("S", st.S, out["S"]); cmp("att_path", st.att_path, out["path"]); cmp("att_omic", st.att_omic, out["omic"])
