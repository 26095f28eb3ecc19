"""A few MCAT train steps (B slides x N patches) for profiling the tail kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, warnings
warnings.filterwarnings("ignore")
from importlib import import_module
pkg = "multimodal-path-omic_b200."
synth = import_module(pkg + "synth"); sp = import_module(pkg + "slidepath"); bpm = import_module(pkg + "bagpass")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = import_module(pkg + "mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES)).to(dev).train()
tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=B)
x = torch.randn((B * N, 1024), device=dev).to(torch.bfloat16)
bag = bpm.PackedBag(x, (N,) * B)
omics = [torch.randn((B, d), device=dev) for d in synth.OMIC_SIZES]
labels = torch.randint(0, 4, (B,), device=dev); censor = torch.randint(0, 2, (B,), device=dev).float()
for _ in range(steps):
    loss, _, _ = tr.step(bag, omics, labels, censor, train=True, seed=7)
torch.cuda.synchronize()
print("loss", loss.mean().item())
