"""SM clock and power while one bag pass runs back to back: `fwd` (MPO_FWD_DEBUG / MPO_FWD_PAIR select the build) or `bwd`."""
import os, sys, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpo_b200
from importlib import import_module
bp = import_module("multimodal-path-omic_b200.bagpass")
B = 32
lengths = (16384,) * B
x = torch.randn(sum(lengths), 1024, device="cuda").bfloat16()
bag = bp.PackedBag(x, lengths)
w = (torch.randn(256, 1024, device="cuda") / 32).bfloat16()
bias = torch.zeros(256, device="cuda"); qk = torch.randn(B, 6, 256, device="cuda") * 0.05
ws = bp.BagWorkspace(bag, save_h=True)
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dpooled = torch.randn(B, 6, 256, device="cuda") * 0.1
gw = torch.zeros(256, 1024, device="cuda"); gb = torch.zeros(256, device="cuda")
bp.bag_forward(bag, w, bias, qk, ws)
fn = (lambda: bp.bag_forward(bag, w, bias, qk, ws)) if which == "fwd" else (lambda: bp.bag_backward(bag, ws, dpooled, qk, gw, gb))
samples = []
stop = False
def poll():
    while not stop:
        out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
        samples.append(out)
        time.sleep(0.1)
for _ in range(5): fn()
torch.cuda.synchronize()
t = threading.Thread(target=poll); t.start()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
n = 6000
for _ in range(n): fn()
e1.record(); torch.cuda.synchronize()
stop = True; t.join()
ms = e0.elapsed_time(e1) / n
mid = samples[len(samples) // 3:]
print(which + " debug=%s pair=%s: %.2f us/slide over %.1f s; clock/power/cap samples (last two thirds): %s" % (
    os.environ.get("MPO_FWD_DEBUG", "0"), os.environ.get("MPO_FWD_PAIR", "0"), ms * 1e3 / B, ms * n / 1e3, " | ".join(mid[:: max(1, len(mid) // 6)])))
