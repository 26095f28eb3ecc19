import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, warnings
warnings.filterwarnings("ignore")
from helpers import load_ge_case
import mpo_oracle as orc
from importlib import import_module
ge = import_module("multimodal-path-omic_b200.ge_nacagat")
name = sys.argv[1] if len(sys.argv) > 1 else "ge_300"
case = load_ge_case(name)
net = ge.GeneExprNarrowContextualAttentionGateTransformer()
net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
net = net.cuda().eval()
with torch.no_grad():
    Y, att = net(wsi=torch.from_numpy(case["bag"]).cuda())
out = orc.ge_forward_backward(case["state"], case["bag"], None)
A = att["attn"].cpu().numpy().astype(np.float64); Ar = out["attn"]
err = np.abs(A - Ar) / (np.abs(Ar) + 1e-3 * Ar.max())
print("attn max err", err.max(), "at", np.unravel_index(err.argmax(), err.shape))
rows = err.max(axis=1); cols = err.max(axis=0)
print("worst rows", np.argsort(-rows)[:8], rows[np.argsort(-rows)[:8]])
print("worst cols", np.argsort(-cols)[:8], cols[np.argsort(-cols)[:8]])
print("row err by tile", [float(rows[i:i+128].max()) for i in range(0, len(rows), 128)])
print("path err", np.max(np.abs(att["path"].cpu().numpy() - out["path"])) / np.max(np.abs(out["path"])))
