"""2-GPU checks (NCCL), run when the box has at least two devices (`gpurun --gpus 2 -- python -m pytest tests -m gpu`):
  * one bag sharded by patch range over two ranks + cross-GPU log-sum-exp combine == the reference outputs
    (golden fixture generated from the unmodified reference, BASELINE config 5);
  * data-parallel gradients: two ranks x 2 slides, one all-reduce == one rank x 4 slides;
  * the split-graph step with the bucketed, overlapped all-reduce == the eager step with one all-reduce."""
import os
import sys
from importlib import import_module

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _pkg(name):
    return import_module("multimodal-path-omic_b200." + name)


def _build(case, dev):
    synth = _pkg("synth")
    cls = _pkg("mcat").MultimodalCoAttentionTransformer if case["model"] == "mcat" else \
        _pkg("nacagat").NarrowContextualAttentionGateTransformer
    net = cls(omic_sizes=list(synth.OMIC_SIZES), fusion=case["fusion"])
    net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
    return net.to(dev).eval()


def _worker(rank, world, port, q):
    import warnings
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dp = _pkg("dp"); sp = _pkg("slidepath"); bpm = _pkg("bagpass"); synth = _pkg("synth")
    # ---- 1) patch-range sharded inference
    case = load_case("mcat_concat_4096")
    net = _build(case, dev)
    a, b = dp.patch_range(case["n"], rank, world)
    wsi = torch.from_numpy(case["bag"][a:b]).to(dev)
    omics = [torch.from_numpy(o).to(dev) for o in case["omics"]]
    hz, S, Y, amap = dp.sharded_inference(net, wsi, omics)
    res = dict(rank=rank, hazards=hz.cpu().numpy(), amap=amap.cpu().numpy(), rows=(a, b))
    full = dp.gather_attention_map(amap, case["n"])                      # the map export path
    res["amap_full"] = None if full is None else full.cpu().numpy()
    # a bag shorter than one tile per rank: rank 1 holds ZERO rows and still takes part in the combine
    small = load_case("mcat_concat_128")
    net_s = _build(small, dev)
    a2, b2 = dp.patch_range(small["n"], rank, world)
    hz2, _, _, amap2 = dp.sharded_inference(net_s, torch.from_numpy(small["bag"][a2:b2]).to(dev).reshape(-1, 1024),
                                            [torch.from_numpy(o).to(dev) for o in small["omics"]])
    res["small"] = dict(hazards=hz2.cpu().numpy(), cols=int(amap2.shape[1]), rows=(a2, b2))
    # ---- 2) data-parallel gradients
    lens = [300, 517, 129, 1000]
    slides = [synth.make_slide(300 + i, n) for i, n in enumerate(lens)]
    mine = dp.slide_shard(len(slides), rank, world)
    tr = sp.BatchTrainer(net, loss="nll", grad_acc_step=len(slides))
    tr.zero_grad()
    pb = bpm.PackedBag.from_slides([torch.from_numpy(slides[i][0]).to(dev) for i in mine])
    om = [torch.stack([torch.from_numpy(slides[i][1][j]) for i in mine]).to(dev) for j in range(6)]
    labels = torch.tensor([slides[i][2] for i in mine], dtype=torch.int64, device=dev)
    cens = torch.tensor([slides[i][3] for i in mine], dtype=torch.float32, device=dev)
    tr.step(pb, om, labels, cens, train=False)
    dp.all_reduce_gradients(tr.flat_grad)
    res["grad_dp"] = tr.flat_grad.cpu().numpy()
    if rank == 0:
        tr.zero_grad()
        pb = bpm.PackedBag.from_slides([torch.from_numpy(s[0]).to(dev) for s in slides])
        om = [torch.stack([torch.from_numpy(s[1][j]) for s in slides]).to(dev) for j in range(6)]
        labels = torch.tensor([s[2] for s in slides], dtype=torch.int64, device=dev)
        cens = torch.tensor([s[3] for s in slides], dtype=torch.float32, device=dev)
        tr.step(pb, om, labels, cens, train=False)
        res["grad_one"] = tr.flat_grad.cpu().numpy()
    # ---- 3) the step captured as two CUDA graphs with the post-stage gradient bucket all-reduced next to the bag
    #         backward pass (bench.py at N > 1) == the eager step followed by one all-reduce
    tr.zero_grad()
    pb = bpm.PackedBag.from_slides([torch.from_numpy(slides[i][0]).to(dev) for i in mine])
    om = [torch.stack([torch.from_numpy(slides[i][1][j]) for i in mine]).to(dev) for j in range(6)]
    labels = torch.tensor([slides[i][2] for i in mine], dtype=torch.int64, device=dev)
    cens = torch.tensor([slides[i][3] for i in mine], dtype=torch.float32, device=dev)
    gstep = tr.capture(pb, om, labels, cens, train=False, split=True)
    off = tr.post_bucket_offset()
    tr.zero_grad()
    gstep.replay_first()
    w1 = dist.all_reduce(tr.flat_grad[off:], async_op=True)
    gstep.replay_second()
    w2 = dist.all_reduce(tr.flat_grad[:off], async_op=True)
    w1.wait(); w2.wait()
    torch.cuda.synchronize()
    res["grad_split"] = tr.flat_grad.cpu().numpy()
    res["bucket_offset"] = int(off)
    torch.cuda.synchronize()
    q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharded_inference_and_dp_gradients():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    case = load_case("mcat_concat_4096")
    g = case["gold"]
    for r in got:
        assert np.max(np.abs(r["hazards"] - g["hazards"]) / np.abs(g["hazards"])) < 1e-3
    amap = np.concatenate([r["amap"] for r in got], axis=1).astype(np.float64)
    ref = g["coattn"].astype(np.float64)
    assert amap.shape == ref.shape
    assert np.max(np.abs(amap - ref) / (np.abs(ref) + 1e-3 * ref.max())) < 1e-3
    assert got[1]["amap_full"] is None and np.array_equal(got[0]["amap_full"], np.concatenate([r["amap"] for r in got], axis=1))
    gs_ = load_case("mcat_concat_128")["gold"]
    assert got[1]["small"]["rows"] == (128, 128) and got[1]["small"]["cols"] == 0 and got[0]["small"]["cols"] == 128
    for r in got:
        assert np.max(np.abs(r["small"]["hazards"] - gs_["hazards"]) / np.abs(gs_["hazards"])) < 1e-3
    g1, gd = got[0]["grad_one"].astype(np.float64), got[0]["grad_dp"].astype(np.float64)
    assert np.linalg.norm(gd - g1) / np.linalg.norm(g1) < 2e-3
    assert np.allclose(got[0]["grad_dp"], got[1]["grad_dp"])
    gs = got[0]["grad_split"].astype(np.float64)
    assert 0 < got[0]["bucket_offset"] < gs.size
    assert np.linalg.norm(gs - gd) / np.linalg.norm(gd) < 1e-4
    assert np.allclose(got[0]["grad_split"], got[1]["grad_split"])
