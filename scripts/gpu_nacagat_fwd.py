"""GPU bring-up of the NaCAGaT forward path against the golden fixtures (forward outputs only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, warnings
warnings.filterwarnings("ignore")
from helpers import golden_cases, load_case
from importlib import import_module
synth = import_module("multimodal-path-omic_b200.synth")
cls = import_module("multimodal-path-omic_b200.nacagat").NarrowContextualAttentionGateTransformer
ok = True
for name in golden_cases():
    if not name.startswith("nacagat"):
        continue
    case = load_case(name)
    net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=case["fusion"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    net = net.cuda().eval()
    wsi = torch.from_numpy(case["bag"]).cuda(); omics = [torch.from_numpy(o).cuda() for o in case["omics"]]
    for mode in ("nograd", "grad"):
        ctx = torch.no_grad() if mode == "nograd" else torch.enable_grad()
        with ctx:
            hazards, S, Y, att = net(wsi=wsi, omics=omics)
        g = case["gold"]
        eh = float(np.max(np.abs(hazards.detach().cpu().numpy() - g["hazards"]) / np.abs(g["hazards"])))
        A, Aref = att["coattn"].detach().cpu().numpy().astype(np.float64), g["coattn"].astype(np.float64)
        ea = float(np.max(np.abs(A - Aref) / (np.abs(Aref) + 1e-3 * Aref.max())))
        ep = float(np.max(np.abs(att["path"].cpu().numpy() - g["path"])) / np.max(np.abs(g["path"])))
        print(f"{name} [{mode}]: hazards {eh:.2e}  coattn {ea:.2e}  path {ep:.2e}")
        ok = ok and eh < 1e-3 and ea < 1e-3 and ep < 1e-3
print("NACAGAT FWD", "PASS" if ok else "FAIL")
