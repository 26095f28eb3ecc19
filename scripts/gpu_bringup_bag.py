"""GPU bring-up check of the bag stage against plain torch ops (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpo_b200 as mpo
from importlib import import_module
bp = import_module("multimodal-path-omic_b200.bagpass")

torch.manual_seed(0)
dev = "cuda"

def ref_fwd(x, w, b, qk):
    # x bf16 [N,1024], w bf16 [256,1024]; fp32 math on the bf16 values
    h = torch.relu(x.float() @ w.float().t() + b)
    s = qk @ h.t()                       # [6,N]
    lse = torch.logsumexp(s, dim=1)
    a = torch.softmax(s, dim=1)
    hb = h.half().float()
    pooled = a @ hb
    return h, hb, s, lse, a, pooled

def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-6)).item()

def run(lengths, tag):
    B = len(lengths)
    slides = [torch.randn(n, 1024, device=dev) for n in lengths]
    bag = bp.PackedBag.from_slides(slides)
    w = (torch.randn(256, 1024, device=dev) / 32).bfloat16()
    bias = torch.randn(256, device=dev) * 0.1
    qk = torch.randn(B, 6, 256, device=dev) * 0.05
    ws = bp.BagWorkspace(bag, save_h=True)
    bp.bag_forward(bag, w, bias, qk, ws)
    torch.cuda.synchronize()
    amap = bp.attention_map(bag, ws)
    torch.cuda.synchronize()
    worst = 0.0
    for b in range(B):
        r0, r1 = bag.slide_rows(b)
        h, hb, s, lse, a, pooled = ref_fwd(bag.x[r0:r1], w, bias, qk[b])
        e = dict(
            scores=(ws.scores[:, r0:r1] - s).abs().max().item(),
            lse=(ws.lse[b] - lse).abs().max().item(),
            pooled=rel(ws.pooled[b], pooled),
            hsaved=rel(ws.h_saved[r0:r1].float(), hb),
            amap=((amap[:, r0:r1] - a).abs() / (a.abs() + 1e-12)).max().item(),
        )
        worst = max(worst, *e.values())
        if b < 3 or max(e.values()) > 1e-2:
            print(f"[{tag}] fwd slide {b} N={lengths[b]}: " + " ".join(f"{k}={v:.3e}" for k, v in e.items()))
    # backward
    dpooled = torch.randn(B, 6, 256, device=dev) * 0.1
    gw = torch.zeros(256, 1024, device=dev)
    gb = torch.zeros(256, device=dev)
    dqk = bp.bag_backward(bag, ws, dpooled, qk, gw, gb)
    torch.cuda.synchronize()
    # torch reference via autograd on the same function with h rounded to bf16 (straight-through)
    gw_ref = torch.zeros_like(gw); gb_ref = torch.zeros_like(gb)
    for b in range(B):
        r0, r1 = bag.slide_rows(b)
        x = bag.x[r0:r1].float()
        wf = w.float().clone().requires_grad_(True)
        bf = bias.clone().requires_grad_(True)
        q = qk[b].clone().requires_grad_(True)
        z = x @ wf.t() + bf
        h = torch.relu(z)
        hb = h + (h.half().float() - h).detach()
        s = q @ hb.t()
        a = torch.softmax(s, dim=1)
        pooled = a @ hb
        (pooled * dpooled[b]).sum().backward()
        gw_ref += wf.grad; gb_ref += bf.grad
        e_dqk = rel(dqk[b], q.grad) if lengths[b] > 1 else 0.0
        worst = max(worst, e_dqk)
        if b < 3 or e_dqk > 1e-2:
            print(f"[{tag}] bwd slide {b}: dqk rel={e_dqk:.3e}")
    e_gw, e_gb = rel(gw, gw_ref), rel(gb, gb_ref)
    worst = max(worst, e_gw, e_gb)
    print(f"[{tag}] bwd: dW_H rel={e_gw:.3e} db_H rel={e_gb:.3e}   worst={worst:.3e}")
    return worst

ok = True
for lengths, tag in [((128,), "one-tile"), ((300,), "ragged-1"), ((1, 129, 77, 512), "ragged-4"), ((4096,), "4k"), ((16384, 5000), "16k+5k")]:
    try:
        w = run(lengths, tag)
        ok = ok and w < 2e-2
    except Exception as ex:
        ok = False
        print(f"[{tag}] EXCEPTION {type(ex).__name__}: {ex}")
        break

# timing of the bag stage alone: 8 slides x 16384 patches
if ok:
    lengths = (16384,) * 8
    x = torch.randn(sum(lengths), 1024, device=dev).bfloat16()
    bag = bp.PackedBag(x, lengths)
    w = (torch.randn(256, 1024, device=dev) / 32).bfloat16()
    bias = torch.zeros(256, device=dev); qk = torch.randn(8, 6, 256, device=dev) * 0.05
    ws = bp.BagWorkspace(bag, save_h=True)
    dpooled = torch.randn(8, 6, 256, device=dev) * 0.1
    gw = torch.zeros(256, 1024, device=dev); gb = torch.zeros(256, device=dev)
    for name, fn in [("fwd", lambda: bp.bag_forward(bag, w, bias, qk, ws)),
                     ("bwd", lambda: bp.bag_backward(bag, ws, dpooled, qk, gw, gb))]:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = sum(lengths) * 2048 / (ms * 1e-3) / 1e9
        print(f"timing {name}: {ms*1e3/8:.1f} us/slide  ({gbs:.0f} GB/s algorithmic)")
print("BRINGUP", "PASS" if ok else "FAIL")
