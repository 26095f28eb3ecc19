"""GE-NaCAGaT train step (forward + CE loss + backward) timing at N patches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, warnings
warnings.filterwarnings("ignore")
from importlib import import_module
ge = import_module("multimodal-path-omic_b200.ge_nacagat")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
net = ge.GeneExprNarrowContextualAttentionGateTransformer().cuda().train()
wsi = torch.randn(N, 1024, device="cuda").bfloat16()
label = torch.tensor([1], device="cuda")
def step():
    Y, att = net(wsi=wsi)
    loss = ge.ge_cross_entropy(Y, label)
    loss.backward()
    return loss
for _ in range(2):
    step(); net.zero_grad()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3 if N > 8192 else 10
e0.record()
for _ in range(reps):
    l = step(); net.zero_grad()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"GE-NaCAGaT N={N}: {ms:.1f} ms/slide-step ({1e3/ms:.2f} slides/s), loss {l.item():.4f}, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
