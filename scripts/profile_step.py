"""One eager train step of B slides bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")
import torch
from importlib import import_module
pkg = "multimodal-path-omic_b200."
synth = import_module(pkg + "synth"); sp = import_module(pkg + "slidepath"); bpm = import_module(pkg + "bagpass")
model = sys.argv[1] if len(sys.argv) > 1 else "mcat"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
cls = (import_module(pkg + "mcat").MultimodalCoAttentionTransformer if model == "mcat"
       else import_module(pkg + "nacagat").NarrowContextualAttentionGateTransformer)
torch.manual_seed(0)
net = cls(omic_sizes=list(synth.OMIC_SIZES)).cuda().train()
tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=B)
x = torch.randn(B * N, 1024, device="cuda").bfloat16()
bag = bpm.PackedBag(x, (N,) * B)
omics = [torch.randn(B, d, device="cuda") for d in synth.OMIC_SIZES]
labels = torch.randint(0, 4, (B,), device="cuda"); censor = torch.randint(0, 2, (B,), device="cuda").float()
for _ in range(3):
    tr.step(bag, omics, labels, censor, train=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(bag, omics, labels, censor, train=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step: model", model, "B", B, "N", N)
