"""MCAT bag backward alone (32 slides x 16 384 patches): CUDA-event time per pass; run under
`ncu --metrics gpu__time_duration.sum --clock-control none` for the per-kernel split."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from importlib import import_module
bpm = import_module("multimodal-path-omic_b200.bagpass")
B, N = int(os.environ.get("B", 32)), 16384
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.empty((B * N, 1024), dtype=torch.bfloat16, device=dev)
for b in range(B):
    x[b * N:(b + 1) * N] = torch.randn((N, 1024), generator=g, device=dev).to(torch.bfloat16)
bag = bpm.PackedBag(x, (N,) * B)
w = (torch.randn((256, 1024), generator=g, device=dev) / 32).to(torch.bfloat16)
bias = torch.randn(256, generator=g, device=dev) * 0.05
qk = torch.randn((B, 6, 256), generator=g, device=dev) * 0.05
ws = bpm.BagWorkspace(bag, save_h=True)
bpm.bag_forward(bag, w, bias, qk, ws)
dP = torch.randn((B, 6, 256), generator=g, device=dev) * 1e-2
gw = torch.zeros((256, 1024), device=dev); gb = torch.zeros(256, device=dev)
dqk = bpm.bag_backward(bag, ws, dP, qk, gw, gb)
torch.cuda.synchronize()
print("regen", os.environ.get("MPO_BWD_REGEN", "1"), "checksums dW %.6e db %.6e dqk %.6e" % (
    gw.double().norm().item(), gb.double().norm().item(), dqk.double().norm().item()))
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    bpm.bag_backward(bag, ws, dP, qk, gw, gb)
b_.record(); torch.cuda.synchronize()
print("  bag_bwd %.3f ms" % (a.elapsed_time(b_) / reps))
