#!/usr/bin/env python
"""Headline benchmark: slides/s for one train step (forward + loss + backward + Adam) of the slide hot path at
16 384 patches per slide, on N B200s of one node (data-parallel over slides, weak scaling).

    python bench.py --gpus 1 --steps 200 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the unmodified reference modules (oracle/_ref) on the host cores

Prints ONE JSON line on rank 0 (contract in the task description / DESIGN.md section "Measurement").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

N_PATCH = 16384
ALGO_BYTES_PER_PATCH_PASS = 2048          # one bf16 row of 1024 features, read once per pass


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default=os.environ.get("MPO_BENCH_MODEL", "mcat"), choices=["mcat", "nacagat"])
    ap.add_argument("--batch", type=int, default=32, help="slides per GPU per step (= the reference's grad_acc_step)")
    ap.add_argument("--patches", type=int, default=N_PATCH)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary configurations (NaCAGaT, scaled window)")
    ap.add_argument("--no-parity", action="store_true", help="skip the reference-fixture parity gate in front of the timing")
    ap.add_argument("--workload", default="train", choices=["train", "shard200k"],
                    help="train: the headline train step; shard200k: BASELINE config 5, one 200 000-patch bag sharded by "
                         "patch range over the GPUs (inference)")
    ap.add_argument("--window", default="scaled", choices=["scaled", "strict"],
                    help="accumulation window at N > 1: scaled = --batch slides per GPU per optimizer step (global window "
                         "batch x N), strict = the reference's global window of --batch slides split over the GPUs")
    ap.add_argument("--comm", default=os.environ.get("MPO_BENCH_COMM", "peer"), choices=["peer", "nccl"],
                    help="N > 1: peer = reduce-scatter / Adam / all-gather kernels over NVLink peer memory inside the step "
                         "graph (csrc/peer.cu); nccl = bucketed NCCL all-reduce between two graphs + flat Adam")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU (reference)
def host_threads():
    """threads the CPU arm may use: every core this process is allowed on (torchrun exports OMP_NUM_THREADS=1, which
    would otherwise throttle the reference arm at N > 1 -- VERDICT r1 weak #11)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(model, n_patch, budget_s=12.0, min_steps=2, steps=None, warmup=1):
    """One slide fwd+bwd on the host cores; returns (slides/s, steps timed, median s, kind, threads, description).

    kind "reference": the UNMODIFIED reference modules installed under oracle/_ref by __graft_entry__.build()
    (oracle/install_ref.py), in train() mode with the NLL loss, torch.set_num_threads(all cores) -- BASELINE.md section 4.
    kind "port": the numpy restatement (oracle/mpo_oracle.py, eval mode), only when oracle/_ref is not installed."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    threads = host_threads()
    import install_ref
    if install_ref.available():
        n = steps
        if n is None:                                   # size the sample from one probe step
            med, _, _ = install_ref.time_train_step(model, n_patch, steps=1, warmup=1, threads=threads)
            n = int(max(min_steps, min(40, budget_s / max(med, 1e-3))))
            warmup = 0
        med, times, threads = install_ref.time_train_step(model, n_patch, steps=n, warmup=warmup, threads=threads)
        return 1.0 / med, len(times), med, "reference", threads, \
            "unmodified reference modules (oracle/_ref), train() mode, NLL loss, fp32 torch CPU"
    import numpy as np
    import mpo_oracle as orc
    from importlib import import_module
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=threads)
    except Exception:
        pass
    synth = import_module("multimodal-path-omic_b200.synth")
    orc.use_dtype(np.float32)
    shapes = reference_shapes(model)
    state = synth.make_state(shapes, 0, model=model)
    bag, omics, label, censor = synth.make_slide(0, n_patch)
    times = []
    t_end = time.time() + budget_s
    while True:
        t0 = time.time()
        orc.model_forward_backward(state, bag, omics, label, censor, model=model, fusion="concat", loss="nll")
        times.append(time.time() - t0)
        if steps is not None:
            if len(times) >= steps:
                break
        elif len(times) >= min_steps and time.time() > t_end:
            break
    orc.use_dtype(np.float64)
    times.sort()
    med = times[len(times) // 2]
    return 1.0 / med, len(times), med, "port", threads, "oracle port (numpy fp32, eval mode): oracle/_ref not installed"


def reference_shapes(model):
    """state_dict shapes of the drop-in module (identical to the reference's, tests/test_modules.py checks it)."""
    import torch
    from importlib import import_module
    synth = import_module("multimodal-path-omic_b200.synth")
    if model == "mcat":
        cls = import_module("multimodal-path-omic_b200.mcat").MultimodalCoAttentionTransformer
    else:
        cls = import_module("multimodal-path-omic_b200.nacagat").NarrowContextualAttentionGateTransformer
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = cls(omic_sizes=list(synth.OMIC_SIZES))
    return {k: tuple(v.shape) for k, v in net.state_dict().items()}


def run_reference(args, rank):
    if rank != 0:
        return
    warm = max(1, min(args.warmup, 2))
    # each "step" of this arm is ONE slide fwd+bwd (a bounded sample of the 32-slide step of our arm: the rate per
    # slide is the metric); bounded so that the whole run ends within a few minutes
    steps = max(1, min(args.steps, 40))
    rate, n, med, kind, threads, what = cpu_reference_rate(args.model, args.patches, steps=steps, warmup=warm)
    line = {
        "impl": "reference", "metric": "slides/sec (fwd+bwd, 16k patches)", "value": rate, "unit": "slides/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model}_train_step_{args.patches}_patches", "slides_per_step": 1,
                   "note": what + "; each step is one slide fwd+bwd on the host cores (median of %d)" % n},
        "cpu_baseline": {"value": rate, "unit": "slides/s", "cores": threads, "kind": kind,
                         "sample": f"{n} x 1 slide of {args.patches} patches, {what}, {threads} threads"},
        "e2e": {"value": rate, "unit": "slides/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ parity gate
def parity_gate(args, dev, pkg):
    """Before anything is timed: the reference fixture at the benchmarked shape (tests/golden/<model>_concat_16384.npz,
    generated from the unmodified reference) replicated over the B slides of one step and run through the SAME captured
    CUDA-graph step the bench times (eval mode: the fixture has no dropout).  Every slide must reproduce the
    reference's hazards / loss / attention map within 1e-3, and the accumulated gradient (B x 1/B of the slide's) the
    reference's gradient digests within 1e-2.  Raises if not: a fast step with wrong results is not a result."""
    import numpy as np
    import torch
    from importlib import import_module
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import digest_errors, load_case
    synth = import_module(pkg + "synth")
    sp = import_module(pkg + "slidepath")
    bpm = import_module(pkg + "bagpass")
    case = load_case(f"{args.model}_concat_16384")
    cls = import_module(pkg + "mcat").MultimodalCoAttentionTransformer if args.model == "mcat" else \
        import_module(pkg + "nacagat").NarrowContextualAttentionGateTransformer
    net = cls(omic_sizes=list(synth.OMIC_SIZES))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    net = net.to(dev).eval()
    B, N = args.batch, case["n"]
    one = torch.from_numpy(case["bag"]).to(dev).to(torch.bfloat16)
    x = one.repeat(B, 1)
    bag = bpm.PackedBag(x, (N,) * B)
    omics = [torch.from_numpy(o).to(dev).reshape(1, -1).repeat(B, 1).contiguous() for o in case["omics"]]
    labels = torch.full((B,), case["label"], dtype=torch.int64, device=dev)
    censor = torch.full((B,), case["censor"], dtype=torch.float32, device=dev)
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=B)
    g = tr.capture(bag, omics, labels, censor, train=False)
    tr.zero_grad()
    loss, hz, S = g.replay()
    torch.cuda.synchronize()
    gold = case["gold"]
    hz, S, loss = hz.cpu().numpy(), S.cpu().numpy(), loss.cpu().numpy()
    amap = tr.engine.attention_map(g.state).cpu().numpy().astype(np.float64)
    Aref = gold["coattn"].astype(np.float64)
    e_map = 0.0
    for b in (0, B // 2, B - 1):
        e_map = max(e_map, float(np.max(np.abs(amap[:, b * N:(b + 1) * N] - Aref) / (np.abs(Aref) + 1e-3 * Aref.max()))))
    grads = {k: v.detach().cpu().numpy() for k, v in tr.grads.items()}
    worst, details = digest_errors(case, grads)
    out = {"fixture": f"{args.model}_concat_16384 x {B} slides through the captured step (eval mode)",
           "hazards_rel_err": float(np.max(np.abs(hz - gold["hazards"]) / np.abs(gold["hazards"]))),
           "S_rel_err": float(np.max(np.abs(S - gold["S"]) / np.abs(gold["S"]))),
           "loss_abs_err": float(np.max(np.abs(loss - float(gold["loss_nll"])))),
           "coattn_rel_err": e_map, "grad_worst_rel_err": float(worst),
           "tolerances": {"outputs": 1e-3, "gradients": 1e-2}}
    out["ok"] = bool(out["hazards_rel_err"] < 1e-3 and out["S_rel_err"] < 1e-3 and out["loss_abs_err"] < 3e-3
                     and out["coattn_rel_err"] < 1e-3 and out["grad_worst_rel_err"] < 1e-2)
    del g, tr, net, x, bag
    torch.cuda.empty_cache()
    if not out["ok"]:
        raise RuntimeError("bench.py parity gate failed: %s" % json.dumps(out))
    return out


# ------------------------------------------------------------------------------------------------ config 5
def run_shard200k(args, rank, world, dev, pkg, peer_group, n_total=200000, steps=None, quick=False):
    """BASELINE config 5: MCAT inference on ONE 200 000-patch bag sharded by patch range over the GPUs.
    Every rank streams its tile-aligned share of the bag, the partial soft-max states are exchanged and merged by
    mpo_peer_lse_combine (one kernel over NVLink peer memory), the tail is replicated; the whole call is one CUDA-graph
    replay per rank.  Checked in the same run against the unsharded call on rank 0 (hazards and the gathered [6, N] map).
    At N = 1 the same graph runs with a group of one (no exchange partner): the single-GPU reference point."""
    import torch
    import torch.distributed as dist
    from importlib import import_module
    synth = import_module(pkg + "synth")
    dp = import_module(pkg + "dp")
    torch.manual_seed(0)
    net = import_module(pkg + "mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES)).to(dev).eval()
    a, b = dp.patch_range(n_total, rank, world)
    # every rank draws the same bag (same generator seed, chunks of 25 000 rows) and keeps its own patch range
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321)
    wsi = torch.empty((b - a, 1024), dtype=torch.bfloat16, device=dev)
    full = torch.empty((n_total, 1024), dtype=torch.bfloat16, device=dev) if rank == 0 else None
    for c0 in range(0, n_total, 25000):
        c1 = min(n_total, c0 + 25000)
        chunk = torch.randn((c1 - c0, 1024), generator=gen, device=dev).to(torch.bfloat16)
        lo, hi = max(a, c0), min(b, c1)
        if hi > lo:
            wsi[lo - a:hi - a] = chunk[lo - c0:hi - c0]
        if full is not None:
            full[c0:c1] = chunk
    omics = [torch.randn(d, generator=gen, device=dev) for d in synth.OMIC_SIZES]
    if world > 1 and peer_group is None:
        raise RuntimeError("the sharded-inference benchmark runs over the peer-memory kernels: use --comm peer")

    class _Solo:                       # world 1: the "merge" is the identity
        def lse_combine(self, lse_l, pooled_l, lse_out, pooled_out, slot=0):
            return lse_out, pooled_out
    sh = dp.ShardedInference(net, wsi, omics, peer_group if world > 1 else _Solo())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    steps = steps or max(20, args.steps)
    for _ in range(max(3, args.warmup)):
        sh.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        hz, S, Y, amap = sh.replay()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_bag = float(ms.item()) / steps
    # parity inside the run: the unsharded module call on rank 0
    check = None
    full_map = dp.gather_attention_map(amap, n_total) if world > 1 else amap
    if rank == 0:
        with torch.no_grad():
            hz1, _, _, att1 = net(full, omics, inference=True)
        ref_map = att1["coattn"]
        check = {"hazards_rel_err": float(((hz - hz1).abs() / hz1.abs()).max().item()),
                 "map_rel_err": float(((full_map - ref_map).abs() / (ref_map.abs() + 1e-3 * ref_map.max())).max().item())}
        check["ok"] = bool(check["hazards_rel_err"] < 1e-3 and check["map_rel_err"] < 1e-3)
    out = {"ms_per_bag": ms_bag, "bags_per_s": 1e3 / ms_bag, "patches": n_total, "patches_per_gpu": b - a,
           "per_gpu_bag_GBps": (b - a) * 2048 / (ms_bag * 1e-3) / 1e9, "aggregate_bag_GBps": n_total * 2048 / (ms_bag * 1e-3) / 1e9,
           "gpu_launches_per_call": sh.launches_per_replay, "parity_vs_unsharded": check,
           "note": "one CUDA-graph replay per rank: SNN + query fold, bag forward over the rank's patch range, "
                   "mpo_peer_lse_combine over NVLink peer memory, replicated tail, map slice; CUDA events, max over ranks"}
    if quick:
        return out
    return {"metric": "bags/sec (MCAT inference, 200k patches sharded by patch range)", "value": 1e3 / ms_bag, "unit": "bags/s",
            "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup), "ms_per_step": ms_bag, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "mcat_inference_200000_patches_sharded", "parallelism": f"patch-range x{world}",
                       "l2": "every call streams the shard from HBM (51 MB per GPU at 8 GPUs < 126 MB L2: the bag of the "
                             "previous call may still be L2-resident; the 1-GPU line streams 410 MB)"},
            "gpu_launches": sh.launches_per_replay * steps, "shard200k": out}


# ------------------------------------------------------------------------------------------------ ours
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from importlib import import_module
    import warnings
    warnings.filterwarnings("ignore")

    pkg = "multimodal-path-omic_b200."
    synth = import_module(pkg + "synth")
    sp = import_module(pkg + "slidepath")
    bpm = import_module(pkg + "bagpass")
    lib = import_module(pkg + "_lib")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one process per GPU: run on the cores (and first-touch the pinned staging memory on the NUMA node) next to this GPU
    numa_cpus = None
    if world > 1 and os.environ.get("MPO_BENCH_NUMA_BIND", "1") == "1":
        numa_cpus = import_module(pkg + "ingest").bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    parity = None if args.no_parity else parity_gate(args, dev, pkg)
    peer_group = None
    if world > 1 and args.comm == "peer":
        peer_group = import_module(pkg + "peer").PeerGroup(dev)
    if args.workload == "shard200k":
        line = run_shard200k(args, rank, world, dev, pkg, peer_group)
        if rank == 0:
            emit(line)
        if peer_group is not None:
            peer_group.close()
        if world > 1:
            dist.destroy_process_group()
        return

    if args.model == "mcat":
        cls = import_module(pkg + "mcat").MultimodalCoAttentionTransformer
    else:
        cls = import_module(pkg + "nacagat").NarrowContextualAttentionGateTransformer
    torch.manual_seed(0)
    net = cls(omic_sizes=list(synth.OMIC_SIZES)).to(dev)
    net.train()
    B, N = args.batch, args.patches
    if world > 1 and args.window == "strict":
        B = max(1, args.batch // world)          # the reference's global window (grad_acc_step slides) split over the GPUs
    trainer = sp.BatchTrainer(net, loss="nll", grad_acc_step=B * world)
    # Adam(lr 2e-4, wd 1e-5): reference mcat/main.py:298, config.yaml
    trainer.use_flat_adam(lr=2e-4, weight_decay=1e-5, peer=peer_group)

    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    lengths = (N,) * B
    x = torch.empty((B * N, 1024), dtype=torch.bfloat16, device=dev)
    for b in range(B):      # chunked: avoids a 4 GB fp32 temporary
        x[b * N:(b + 1) * N] = torch.randn((N, 1024), generator=gen, device=dev, dtype=torch.float32).to(torch.bfloat16)
    bag = bpm.PackedBag(x, lengths)
    omics = [torch.randn((B, d), generator=gen, device=dev) for d in synth.OMIC_SIZES]
    labels = torch.randint(0, 4, (B,), generator=gen, device=dev, dtype=torch.int64)
    censor = torch.randint(0, 2, (B,), generator=gen, device=dev).to(torch.float32)

    # the whole step replays as one CUDA graph; on one GPU the Adam update rides in it, with several GPUs the
    # gradient all-reduce sits between the graph and the (then eager, two-launch) Adam update
    # (N > 1: two graphs, so that the all-reduce of the post-stage gradient bucket -- ~70 % of the 16.6 MB, final once
    # the tail's path kernel and its weight-gradient kernel have run -- overlaps the bag backward pass)
    # opt-in (MPO_BENCH_NCCL_IN_GRAPH=1): measured 1.174 vs 1.191 ms per step at N = 2, but the process then hangs in
    # the process-group teardown with the captured collectives alive, so the default keeps NCCL outside the graphs
    one_graph = world > 1 and os.environ.get("MPO_BENCH_NCCL_IN_GRAPH", "0") == "1"
    graphed = None
    if peer_group is not None:
        # the whole data-parallel step -- communication and optimizer included -- is ONE graph of plain kernels
        graphed = trainer.capture(bag, omics, labels, censor, train=True, peer_step=True,
                                  peer_overlap=os.environ.get("MPO_BENCH_PEER_OVERLAP", "1") == "1")
        one_graph = True
    elif one_graph:
        # the bucketed all-reduce and the Adam update inside the captured step (NCCL collectives are graph-capturable)
        try:
            graphed = trainer.capture(bag, omics, labels, censor, train=True, with_adam=True, allreduce=True)
        except Exception as exc:          # fall back to two graphs with the collectives between them
            print("bench: NCCL-in-graph capture failed (%s); using the split-graph step" % (exc,), file=sys.stderr)
            one_graph = False
            torch.cuda.synchronize()
    if graphed is None:
        graphed = trainer.capture(bag, omics, labels, censor, train=True, with_adam=(world == 1), split=(world > 1))
    post_off = trainer.post_bucket_offset() if world > 1 else 0

    def one_step():
        if world == 1 or one_graph:
            loss, _, _ = graphed.replay()
            return loss
        graphed.replay_first()
        w1 = dist.all_reduce(trainer.flat_grad[post_off:], async_op=True)      # NCCL stream, next to the bag backward
        loss, _, _ = graphed.replay_second()
        w2 = dist.all_reduce(trainer.flat_grad[:post_off], async_op=True)      # H, SNN and co-attention in-projection
        w1.wait()
        w2.wait()
        trainer.adam_step(zero_grad=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        one_step()
    barrier()
    lib.lib().mpo_launch_count(1)
    graphed.replays = 0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    # kernels of libmpo_b200.so inside the timed region: the graph's kernel nodes x replays (+ any eager launches)
    launches = int(lib.lib().mpo_launch_count(0)) + graphed.replays * graphed.launches_per_replay
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # the timed region is tens of milliseconds; keep the SAME step running (untimed, same count on every rank) for
    # about a second so that nvidia-smi samples the clocks under this load
    for _ in range(int(min(5000, max(1, 1000.0 * args.steps / max(ms, 1e-3))))):
        one_step()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- per-stage device times and the roofline of the dominant kernel (rank 0, separate pass)
    stages, roof = {}, None
    if rank == 0:
        eng = trainer.engine
        st = graphed.state
        dpooled = torch.randn((B, 6, 256), device=dev) * 1e-3
        gw = torch.zeros((256, 1024), device=dev)
        gb = torch.zeros(256, device=dev)
        P = dict(net.named_parameters())

        def t_stage(fn, reps=5):
            fn(); torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b_.record(); torch.cuda.synchronize()
            return a.elapsed_time(b_) / reps

        if args.model == "mcat":
            stages["bag_fwd_ms"] = t_stage(lambda: bpm.bag_forward(bag, eng._w_bf16, P["H.0.bias"].detach(), st.qk,
                                                                  st.bag_ws, seed=1, drop_p=st.drop_p))
            stages["bag_bwd_ms"] = t_stage(lambda: bpm.bag_backward(bag, st.bag_ws, dpooled, st.qk, gw, gb,
                                                                   drop_p=st.drop_p))
        else:
            # NaCAGaT: projection pass + gate pass forward; the whole mpo_bag_bwd_nacagat chain backward
            bk = P["co_attention.in_proj_bias"].detach()[256:512]

            def nac_fwd():
                bpm.bag_project(bag, eng._w_bf16, P["H.0.bias"].detach(), st.qk, st.bag_ws, seed=1, drop_p=st.drop_p)
                bpm.bag_gate_forward(bag, eng._wk_f16, bk, st.qp, st.kc, st.bag_ws, seed=1, attn_drop_p=st.attn_p)

            def nac_bwd():
                st.dpooled.copy_(dpooled)
                eng.bag_backward_only(trainer.model, st)
            stages["bag_fwd_ms"] = t_stage(nac_fwd)
            stages["bag_bwd_ms"] = t_stage(nac_bwd)
        stages["step_ms"] = ms / args.steps
        stages["tail_and_rest_ms"] = max(0.0, stages["step_ms"] - stages["bag_fwd_ms"] - stages["bag_bwd_ms"])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        algo_bytes = B * N * ALGO_BYTES_PER_PATCH_PASS
        dom = "bag_fwd" if stages["bag_fwd_ms"] >= stages["bag_bwd_ms"] else "bag_bwd"
        dom_ms = stages[dom + "_ms"]
        ach = algo_bytes / (dom_ms * 1e-3) / 1e9
        # DRAM traffic of the same launch (B slides x N patches) from the committed `ncu --set full` capture
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(f"{args.model}_{dom}_B{B}_N{N}")
            if ent:
                traffic = float(ent["dram_bytes_per_launch"])
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom + (" (tcgen05 projection + fused co-attention"
                                                   + (", gate pass" if args.model != "mcat" else "") + ")"
                                                   if dom == "bag_fwd" else " (tcgen05 row-expand dz stage + split-K dW GEMM)"),
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes,
                "whole_step_achieved": 2 * algo_bytes / (stages["step_ms"] * 1e-3) / 1e9,
                "whole_step_frac": 2 * algo_bytes / (stages["step_ms"] * 1e-3) / 1e9 / peak}

    # ---- end to end: host-resident pinned bags, H2D inside the timed region, loss read back every step
    e2e = None
    if not args.no_e2e:
        nset = 2
        host_x = [torch.empty((B * N, 1024), dtype=torch.bfloat16).pin_memory() for _ in range(nset)]
        for hx in host_x:
            hx.copy_(x.cpu())
        host_om = [[o.cpu().pin_memory() for o in omics] for _ in range(nset)]
        host_lab = labels.cpu().pin_memory(); host_cen = censor.cpu().pin_memory()
        dev_x = [torch.empty_like(x) for _ in range(2)]
        dev_bags = [bpm.PackedBag(dx, lengths) for dx in dev_x]
        dev_om = [[o.clone() for o in omics] for _ in range(2)]
        dev_lab = [labels.clone() for _ in range(2)]
        dev_cen = [censor.clone() for _ in range(2)]
        for dx in dev_x:
            dx.copy_(x)
        steps_g = [trainer.capture(dev_bags[i], dev_om[i], dev_lab[i], dev_cen[i], train=True, with_adam=(world == 1),
                                   peer_step=peer_group is not None) for i in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.empty(B, dtype=torch.float32).pin_memory()

        def stage_in(k):
            slot = k % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[slot])
                dev_x[slot].copy_(host_x[k % nset], non_blocking=True)
                for d_, h_ in zip(dev_om[slot], host_om[k % nset]):
                    d_.copy_(h_, non_blocking=True)
                dev_lab[slot].copy_(host_lab, non_blocking=True)
                dev_cen[slot].copy_(host_cen, non_blocking=True)
                ready[slot].record(copy_stream)
            return None

        def e2e_loop(k_steps):
            stage_in(0)
            for k in range(k_steps):
                slot = k % 2
                if k + 1 < k_steps:
                    stage_in(k + 1)
                torch.cuda.current_stream().wait_event(ready[slot])
                loss, _, _ = steps_g[slot].replay()
                freed[slot].record(torch.cuda.current_stream())
                if world > 1 and peer_group is None:
                    dist.all_reduce(trainer.flat_grad)
                    trainer.adam_step(zero_grad=True)
                loss_host.copy_(loss, non_blocking=True)
                torch.cuda.current_stream().synchronize()      # the caller reads the loss every step (main.py:49)

        for f in freed:
            f.record(torch.cuda.current_stream())
        e2e_steps = max(2, min(args.steps, 6))
        e2e_loop(2)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(e2e_steps)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        h2d = B * N * 1024 * 2 + sum(B * d * 4 for d in synth.OMIC_SIZES) + B * 8 + B * 4
        e2e = {"value": world * B * e2e_steps / float(dt.item()), "unit": "slides/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": B * 4, "steps": e2e_steps,
               "note": "pinned host bf16 bags, double-buffered H2D on a copy stream, loss read back every step",
               "cpu_affinity_rank0": (None if numa_cpus is None else "%d cores next to GPU %d" % (len(numa_cpus), local_rank))}

    # ---- secondary configurations measured in the same run (N = 1 only): BASELINE config 2 (NaCAGaT train step) and
    # the scaled accumulation window of SURVEY 8e (128 slides per optimizer step instead of the reference's 32)
    also = None
    if world == 1 and not args.no_also:
        del graphed
        torch.cuda.empty_cache()

        def quick_rate(model_name, Bq, steps):
            mcls = import_module(pkg + "mcat").MultimodalCoAttentionTransformer if model_name == "mcat" else \
                import_module(pkg + "nacagat").NarrowContextualAttentionGateTransformer
            torch.manual_seed(0)
            qnet = mcls(omic_sizes=list(synth.OMIC_SIZES)).to(dev).train()
            qtr = sp.BatchTrainer(qnet, loss="nll", grad_acc_step=Bq)
            qtr.use_flat_adam(lr=2e-4, weight_decay=1e-5)
            if Bq == B:
                qx = x
            else:
                qx = torch.empty((Bq * N, 1024), dtype=torch.bfloat16, device=dev)
                for b in range(Bq):
                    qx[b * N:(b + 1) * N] = x[(b % B) * N:(b % B + 1) * N]
            qbag = bpm.PackedBag(qx, (N,) * Bq)
            qom = [torch.randn((Bq, d), generator=gen, device=dev) for d in synth.OMIC_SIZES]
            qlab = torch.randint(0, 4, (Bq,), generator=gen, device=dev, dtype=torch.int64)
            qcen = torch.randint(0, 2, (Bq,), generator=gen, device=dev).to(torch.float32)
            qg = qtr.capture(qbag, qom, qlab, qcen, train=True, with_adam=True)

            def qstep():
                qg.replay()
            for _ in range(3):
                qstep()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(steps):
                qstep()
            b_.record(); torch.cuda.synchronize()
            msq = a.elapsed_time(b_) / steps
            return {"slides_per_s": Bq / (msq * 1e-3), "ms_per_step": msq, "slides_per_step": Bq,
                    "gpu_launches_per_step": qg.launches_per_replay}
        also = {}
        other = "nacagat" if args.model == "mcat" else "mcat"
        also[f"{other}_train_step_{N}_patches_B{B}"] = quick_rate(other, B, max(5, args.steps))
        also[f"{args.model}_scaled_window_B{4 * B}"] = quick_rate(args.model, 4 * B, max(3, args.steps // 2))
        torch.cuda.empty_cache()
        # BASELINE config 4a: GE-NaCAGaT train step (forward + the driver's cross-entropy + backward), one slide
        ge = import_module(pkg + "ge_nacagat")
        torch.manual_seed(0)
        gnet = ge.GeneExprNarrowContextualAttentionGateTransformer().to(dev).train()
        gx = x[:N]
        glab = torch.tensor([1], device=dev)

        def gstep():
            Yg, _ = gnet(wsi=gx)
            ge.ge_cross_entropy(Yg, glab).backward()
            gnet.zero_grad()
        gstep(); torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            gstep()
        b_.record(); torch.cuda.synchronize()
        msg = a.elapsed_time(b_) / 3
        also[f"ge_nacagat_train_step_{N}_patches_B1"] = {
            "slides_per_s": 1e3 / msg, "ms_per_step": msg, "slides_per_step": 1,
            "note": "attention products and N-token weight gradients on tcgen05 (bf16 hi/lo operand pairs, fp32 "
                    "accumulation; csrc/tc_gemm.cu), N x N attention materialised (DESIGN.md 4.5)"}
        del gnet
        torch.cuda.empty_cache()
        # BASELINE config 5: MCAT inference on one 200 000-patch bag.  One GPU streams the whole bag here; under the
        # 8-way patch-range sharding of dp.sharded_inference each GPU streams 25 000 patches (second entry) and the
        # combine adds one all-gather of 6 x 257 floats.
        inet = import_module(pkg + "mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES)).to(dev).eval()
        iom = [torch.randn(d, generator=gen, device=dev) for d in synth.OMIC_SIZES]
        for tag, n_inf in (("mcat_inference_200000_patches_1gpu", 200000), ("mcat_inference_25000_patch_shard", 25000)):
            n_inf = min(n_inf, B * N)
            wsi = x[:n_inf]
            with torch.no_grad():
                for _ in range(2):
                    inet(wsi, iom, inference=True)
                torch.cuda.synchronize()
                a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    inet(wsi, iom, inference=True)
                b_.record(); torch.cuda.synchronize()
            msi = a.elapsed_time(b_) / 5
            also[tag] = {"ms_per_bag": msi, "patches": n_inf, "bag_GBps": n_inf * 2048 / (msi * 1e-3) / 1e9,
                         "note": "eager module call (forward + attention map), CUDA events"}
        del inet
        torch.cuda.empty_cache()
        also["mcat_inference_200000_patches_graphed_1gpu"] = run_shard200k(args, 0, 1, dev, pkg, None, quick=True)
        inet = import_module(pkg + "mcat").MultimodalCoAttentionTransformer(omic_sizes=list(synth.OMIC_SIZES)).to(dev)
        # the drop-in path itself: the reference's per-slide loop (models/mcat/main.py:39-70) calling the module and
        # the loss one slide at a time through torch.autograd (host overhead included: wall clock around 20 calls)
        inet.train()
        lfn = import_module(pkg + "loss").NegativeLogLikelihoodSurvivalLoss()
        wsi1 = x[:N]
        lab1, cen1 = torch.tensor([[1]], device=dev), torch.tensor([0.0], device=dev)

        def slide_step():
            hz_, S_, _, _ = inet(wsi1, iom)
            lfn(hz_, S_, lab1, cen1).backward()
        for _ in range(3):
            slide_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            slide_step()
        torch.cuda.synchronize()
        ms_slide = (time.perf_counter() - t0) / 20 * 1e3
        also[f"mcat_module_api_per_slide_{N}_patches"] = {
            "ms_per_slide_step": ms_slide, "slides_per_s": 1e3 / ms_slide,
            "note": "model(wsi, omics) + NLL loss + loss.backward() per slide, as the reference's train() loop calls it"}
        del inet
        torch.cuda.empty_cache()

    if world > 1 and not args.no_also:
        # the other regime of SURVEY 8e and BASELINE config 5, measured in the same multi-GPU run (every rank takes part)
        also = {}
        other = "strict" if args.window == "scaled" else "scaled"
        Bo = max(1, args.batch // world) if other == "strict" else args.batch
        if peer_group is not None and Bo <= B:
            torch.manual_seed(0)
            onet = cls(omic_sizes=list(synth.OMIC_SIZES)).to(dev).train()
            otr = sp.BatchTrainer(onet, loss="nll", grad_acc_step=Bo * world)
            otr.use_flat_adam(lr=2e-4, weight_decay=1e-5, peer=peer_group)
            obag = bpm.PackedBag(x[:Bo * N], (N,) * Bo)
            og = otr.capture(obag, [o[:Bo].contiguous() for o in omics], labels[:Bo].contiguous(), censor[:Bo].contiguous(),
                             train=True, peer_step=True)
            for _ in range(5):
                og.replay()
            barrier()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            nst = max(20, args.steps)
            for _ in range(nst):
                og.replay()
            b_.record()
            barrier()
            mso = torch.tensor([a_.elapsed_time(b_)], device=dev)
            dist.all_reduce(mso, op=dist.ReduceOp.MAX)
            mso = float(mso.item()) / nst
            also[f"{other}_window"] = {
                "slides_per_s": world * Bo / (mso * 1e-3), "ms_per_step": mso, "slides_per_gpu_per_step": Bo,
                "global_slides_per_step": Bo * world,
                "note": ("the reference's global accumulation window of %d slides split over the GPUs (SURVEY 8e 'strict')"
                         % (Bo * world)) if other == "strict" else "scaled window: --batch slides per GPU per optimizer step"}
            del og, otr, onet
            torch.cuda.empty_cache()
        if peer_group is not None:
            also["mcat_inference_200000_patches_sharded"] = run_shard200k(args, rank, world, dev, pkg, peer_group, quick=True)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, n, med, kind, threads, what = cpu_reference_rate(args.model, N, budget_s=12.0)
        cpu = {"value": rate, "unit": "slides/s", "cores": threads, "kind": kind,
               "sample": f"{n} x 1 slide of {N} patches fwd+bwd, {what}, {threads} threads"}

    if rank == 0:
        line = {
            "metric": "slides/sec (fwd+bwd, 16k patches)", "value": value, "unit": "slides/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model}_train_step_{N}_patches", "slides_per_gpu_per_step": B,
                       "global_slides_per_step": B * world, "patches_per_slide": N, "features": 1024,
                       "parallelism": f"dp{world}", "mode": "train (every dropout layer of the path on: bag embedding, "
                       + ("attention weights, " if args.model != "mcat" else "") + "SNN, encoder layers, pooling heads, rho), "
                       "NLL loss, Adam(lr 2e-4, wd 1e-5) step per batch (mpo_adam_step, in the graph at N=1)",
                       "window": args.window if world > 1 else "single GPU: --batch slides per optimizer step",
                       "comm": None if world == 1 else (
                           "peer: reduce-scatter + Adam + all-gather kernels over NVLink peer memory inside the step graph, the "
                           "post-stage bucket (~2/3 of 16.6 MB) on a graph branch next to the bag backward pass (csrc/peer.cu)"
                           if peer_group is not None else
                           "nccl: fp32 all-reduce in two buckets between two graphs, the post-stage bucket next to the bag "
                           "backward pass, then the flat Adam kernel"),
                       "l2": f"each step streams {B * N * 2048 / 1e9:.2f} GB of bag per GPU (> 126 MB L2), no flush needed"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu,
            "parity": parity,
            "stages": stages, "also": also,
        }
        emit(line)
    if peer_group is not None:
        barrier()
        peer_group.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # the contract is ONE JSON line on stdout: libraries that write banners to fd 1 (NCCL prints its version there)
    # are sent to stderr; the result line goes to the saved descriptor
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)]
        cmd += sys.argv[1:]
        sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
