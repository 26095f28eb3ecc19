"""Per-stage device time of one train step: each C-ABI stage captured as its own CUDA graph and replayed."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, warnings
warnings.filterwarnings("ignore")
from importlib import import_module
pkg = "multimodal-path-omic_b200."
synth = import_module(pkg + "synth"); sp = import_module(pkg + "slidepath"); bpm = import_module(pkg + "bagpass")
lib = import_module(pkg + "_lib")
model_name = sys.argv[1] if len(sys.argv) > 1 else "mcat"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
N = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
dev = torch.device("cuda", 0)
cls = import_module(pkg + model_name).MultimodalCoAttentionTransformer if model_name == "mcat" else \
    import_module(pkg + "nacagat").NarrowContextualAttentionGateTransformer
torch.manual_seed(0)
net = cls(omic_sizes=list(synth.OMIC_SIZES)).to(dev).train()
tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=B)
x = torch.randn((B * N, 1024), device=dev, dtype=torch.float32).to(torch.bfloat16) if B * N <= 2**19 else None
if x is None:
    x = torch.empty((B * N, 1024), dtype=torch.bfloat16, device=dev)
    for b in range(B):
        x[b * N:(b + 1) * N] = torch.randn((N, 1024), device=dev).to(torch.bfloat16)
bag = bpm.PackedBag(x, (N,) * B)
omics = [torch.randn((B, d), device=dev) for d in synth.OMIC_SIZES]
labels = torch.randint(0, 4, (B,), device=dev); censor = torch.randint(0, 2, (B,), device=dev).float()
eng = tr.engine
st = eng.alloc_state(tr.model, bag, save_for_backward=True, with_backward_buffers=True)
st.seed_dev = torch.tensor([12345], dtype=torch.int32, device=dev)
for _ in range(2):
    tr._run(st, bag, omics, labels, censor, True, 0)
torch.cuda.synchronize()
io = eng._io(st); m = tr.model
s = lambda: bpm._stream()
P = dict(net.named_parameters())
stages = {
    "tail_pre_fwd": lambda: lib.call("mpo_tail_pre_fwd", ctypes.byref(m), ctypes.byref(io), s()),
    "tail_post_fwd": lambda: lib.call("mpo_tail_post_fwd", ctypes.byref(m), ctypes.byref(io), s()),
    "loss": lambda: lib.call("mpo_surv_loss", tr.kind, bpm._ptr(st.hazards), bpm._ptr(st.S), bpm._ptr(labels), bpm._ptr(censor),
                             ctypes.c_float(tr.alpha), ctypes.c_float(tr.eps), ctypes.c_float(1.0 / B), bpm._ptr(st.loss),
                             bpm._ptr(st.dhz), bpm._ptr(st.dS), B, 4, s()),
    "tail_post_bwd": lambda: lib.call("mpo_tail_post_bwd", ctypes.byref(m), ctypes.byref(io), bpm._ptr(st.dhz), bpm._ptr(st.dS), None, s()),
    "tail_pre_bwd": lambda: lib.call("mpo_tail_pre_bwd", ctypes.byref(m), ctypes.byref(io), s()),
    "whole_step": lambda: tr._run(st, bag, omics, labels, censor, True, 0),
}
opt = torch.optim.Adam(net.parameters(), lr=2e-4, weight_decay=1e-5, fused=True)
res = {}
for name, fn in stages.items():
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    lib.lib().mpo_launch_count(1)
    with torch.cuda.graph(g):
        fn()
    nl = lib.lib().mpo_launch_count(1)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): g.replay()
    e1.record(); torch.cuda.synchronize()
    res[name] = e0.elapsed_time(e1) / 20
    print(f"{name:16s} {res[name]*1e3:8.1f} us   {nl} launches")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    opt.step(); tr.zero_grad()
e1.record(); torch.cuda.synchronize()
print(f"{'adam+zero_grad':16s} {e0.elapsed_time(e1)/20*1e3:8.1f} us")
