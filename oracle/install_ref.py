"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, installed where the GPU box can import them.

TEST / MEASUREMENT INFRASTRUCTURE, not product code: nothing under multimodal-path-omic_b200/ imports oracle/.

The reference (mattiagualtieri/multimodal-path-omic) is a pure-Python PyTorch package with no build system (no
setup.py / pyproject), so `pip install` has nothing to install; this recipe copies the seven files SURVEY.md 8(a) cites,
byte for byte, from /root/reference into oracle/_ref/models/...  oracle/_ref/ is git-ignored (no reference source
enters the history) but not gpurun-ignored, so it travels to the GPU box like the built .so files.  It is used
  * by bench.py's cpu_baseline leg and `bench.py --impl reference` (kind "reference": the reference's own modules in
    train() mode on the host cores), and
  * by tests/golden/make_golden.py's consistency check.
`__graft_entry__.build()` runs install() whenever /root/reference exists (the build container); on the GPU box the
prebuilt copy is used as it is.  models/utils.py imports h5py at module top (not installed, unused on this
path): load() registers an empty stub module for it, the reference files themselves are never edited.
"""
import hashlib
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ["models/mcat/mcat.py", "models/nacagat/nacagat.py", "models/ge_nacagat/ge_nacagat.py", "models/blocks.py",
         "models/fusion.py", "models/loss.py", "models/utils.py"]


def install(src=REF_SRC, dst=REF_DST):
    """copies FILES (and nothing else) and writes MANIFEST.txt with their sha256; returns the manifest lines."""
    if not os.path.isdir(src):
        raise RuntimeError("reference checkout %s not found" % src)
    lines = []
    for rel in FILES:
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), out)
        lines.append("%s  %s" % (hashlib.sha256(open(out, "rb").read()).hexdigest(), rel))
    with open(os.path.join(dst, "MANIFEST.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    return lines


def available(dst=REF_DST):
    return all(os.path.isfile(os.path.join(dst, rel)) for rel in FILES)


def load(dst=REF_DST):
    """imports the installed reference modules; returns a namespace with the model and loss classes."""
    if not available(dst):
        raise RuntimeError("oracle/_ref is not installed (run oracle/install_ref.py where /root/reference exists)")
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    if "models" in sys.modules and not getattr(sys.modules["models"], "__path__", [""])[0].startswith(dst):
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
    if dst not in sys.path:
        sys.path.insert(0, dst)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from models.mcat.mcat import MultimodalCoAttentionTransformer
        from models.nacagat.nacagat import NarrowContextualAttentionGateTransformer
        from models.ge_nacagat.ge_nacagat import GeneExprNarrowContextualAttentionGateTransformer
        from models.loss import NegativeLogLikelihoodSurvivalLoss, CrossEntropySurvivalLoss
    return types.SimpleNamespace(mcat=MultimodalCoAttentionTransformer, nacagat=NarrowContextualAttentionGateTransformer,
                                 ge=GeneExprNarrowContextualAttentionGateTransformer,
                                 nll=NegativeLogLikelihoodSurvivalLoss, ces=CrossEntropySurvivalLoss)


def time_train_step(model="mcat", n_patches=16384, steps=5, warmup=2, threads=None, seed=0):
    """The reference's own train step on the host cores: model.train(), forward, NLL loss / grad_acc_step, backward
    (models/mcat/main.py:39-70), fp32, one synthetic slide of n_patches x 1024 (mcat.py:151-152 style inputs).
    Returns (median seconds per slide-step, list of step times, threads used)."""
    import time
    import torch
    threads = int(threads or os.cpu_count() or 1)
    torch.set_num_threads(threads)          # torchrun exports OMP_NUM_THREADS=1: set the pool size explicitly
    ref = load()
    omic_sizes = [100, 200, 300, 400, 500, 600]
    torch.manual_seed(seed)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = (ref.mcat if model == "mcat" else ref.nacagat)(omic_sizes=omic_sizes)
    net.train()
    wsi = torch.randn(n_patches, 1024).to(torch.bfloat16).to(torch.float32)      # the bf16-representable bag, in fp32
    omics = [torch.randn(d) for d in omic_sizes]
    Y, c = torch.tensor([[1]]), torch.tensor([0.0])
    loss_fn = ref.nll()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        hazards, S, _, _ = net(wsi=wsi, omics=omics)
        loss = loss_fn(hazards, S, Y, c) / 32
        loss.backward()
        dt = time.perf_counter() - t0
        net.zero_grad()
        if i >= warmup:
            times.append(dt)
    times_sorted = sorted(times)
    return times_sorted[len(times_sorted) // 2], times, threads


if __name__ == "__main__":
    for line in install():
        print(line)
    med, times, th = time_train_step(n_patches=4096, steps=3, warmup=1)
    print("reference MCAT train step, 4096 patches: %.1f ms per slide on %d threads" % (med * 1e3, th))
