"""Timing-only probe of the bag kernels (no correctness check): prints us/slide for fwd and bwd."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpo_b200
from importlib import import_module
bp = import_module("multimodal-path-omic_b200.bagpass")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
which = sys.argv[2] if len(sys.argv) > 2 else "fwd,bwd"
dev = "cuda"
lengths = (16384,) * B
x = torch.randn(sum(lengths), 1024, device=dev).bfloat16()
bag = bp.PackedBag(x, lengths)
w = (torch.randn(256, 1024, device=dev) / 32).bfloat16()
bias = torch.zeros(256, device=dev); qk = torch.randn(B, 6, 256, device=dev) * 0.05
save_h = os.environ.get("MPO_TIME_SAVE_H", "1") != "0"
ws = bp.BagWorkspace(bag, save_h=save_h)
dpooled = torch.randn(B, 6, 256, device=dev) * 0.1
gw = torch.zeros(256, 1024, device=dev); gb = torch.zeros(256, device=dev)
fns = {"fwd": lambda: bp.bag_forward(bag, w, bias, qk, ws), "bwd": lambda: bp.bag_backward(bag, ws, dpooled, qk, gw, gb)}
bp.bag_forward(bag, w, bias, qk, ws)
for name in which.split(","):
    fn = fns[name]
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms*1e3/B:.2f} us/slide  ({sum(lengths)*2048/(ms*1e-3)/1e9:.0f} GB/s algorithmic)  env={os.environ.get('MPO_FWD_DEBUG','0')} cluster={os.environ.get('MPO_FWD_CLUSTER','2')} save_h={int(save_h)}")
