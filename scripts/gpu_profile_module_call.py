"""Host-side cost of the drop-in module call (one slide): cProfile of inference and of a train step."""
import os, sys, cProfile, pstats, io as _io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, warnings, time
warnings.filterwarnings("ignore")
from importlib import import_module
pkg = "multimodal-path-omic_b200."
synth = import_module(pkg + "synth")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = import_module(pkg + "mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES)).to(dev)
loss_fn = import_module(pkg + "loss").NegativeLogLikelihoodSurvivalLoss()
x = torch.randn((N, 1024), device=dev).to(torch.bfloat16)
om = [torch.randn(d, device=dev) for d in synth.OMIC_SIZES]
def infer():
    with torch.no_grad():
        net(x, om, inference=True)
def train():
    hz, S, Y, _ = net(x, om)
    l = loss_fn(hz, S, torch.tensor([[1]], device=dev), torch.tensor([0.0], device=dev))
    l.backward()
for name, fn, mode in (("inference", infer, False), ("train fwd+bwd", train, True)):
    net.train(mode)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize()
    ms = torch.cuda.memory_stats()
    print("%s: %.2f ms per call (N = %d)   device allocs so far %d, alloc retries %d" % (
        name, (time.perf_counter() - t0) / 20 * 1e3, N, ms.get("num_device_alloc", -1), ms.get("num_alloc_retries", -1)))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20): fn()
    torch.cuda.synchronize(); pr.disable()
    s = _io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[5:32]))
