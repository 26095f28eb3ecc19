"""Phase timing of the fused tail's path kernels (needs a -DMPO_TAIL_PROF build of libmpo_b200.so: scripts/run_gpu_tailprof.sh).
Thread 0 of CTA 0 (path role, cluster rank 0) accumulates clock64 intervals; prints cycles per pass and phase."""
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
N = 2048
dev = torch.device("cuda", 0)
cls = import_module(pkg + model_name).MultimodalCoAttentionTransformer if model_name == "mcat" else \
    import_module(pkg + "nacagat").NarrowContextualAttentionGateTransformer
torch.manual_seed(0)
net = cls(omic_sizes=list(synth.OMIC_SIZES)).to(dev).train()
tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=B)
x = torch.randn((B * N, 1024), device=dev).to(torch.bfloat16)
bag = bpm.PackedBag(x, (N,) * B)
omics = [torch.randn((B, d), device=dev) for d in synth.OMIC_SIZES]
labels = torch.randint(0, 4, (B,), device=dev); censor = torch.randint(0, 2, (B,), device=dev).float()
st = tr.engine.alloc_state(tr.model, bag, save_for_backward=True, with_backward_buffers=True)
st.seed_dev = torch.tensor([12345], dtype=torch.int32, device=dev)
for _ in range(3):
    tr._run(st, bag, omics, labels, censor, True, 0)
torch.cuda.synchronize()
L = lib.lib()
buf = (ctypes.c_ulonglong * 32)()
L.mpo_tail_prof_read(buf, 1)
reps = 20
for _ in range(reps):
    tr._run(st, bag, omics, labels, censor, True, 0)
L.mpo_tail_prof_read(buf, 1)
names = ["kernel", "gemm_block total", "ring wait + CTA barrier", "chunk issue", "FFMA2 block", "partial store + CTA barrier",
         "cluster barriers", "gemm_block calls"]
for p, pn in ((0, "forward pass"), (1, "loss + backward pass")):
    print("%s %s B=%d: cycles per launch (thread 0 of CTA 0)" % (model_name, pn, B))
    tot = buf[8 * p] / reps
    for i, n in enumerate(names):
        v = buf[8 * p + i] / reps
        print("  %-28s %10.0f  %5.1f %%" % (n, v, 100 * v / tot if i != 7 else 0))
