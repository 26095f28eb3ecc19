"""2-GPU checks of the hand-written NVLink peer-memory collectives (csrc/peer.cu), against NCCL + the single-GPU kernels:
  * mpo_peer_lse_combine == all-gather + mpo_lse_combine, eagerly and from a captured graph (ShardedInference), with the
    reference fixture's hazards / map as the ground truth;
  * mpo_peer_adam_step (reduce-scatter + Adam + all-gather over peer memory, bucketed) == NCCL all-reduce + mpo_adam_step.
Run with `gpurun --gpus 2 -- python -m pytest tests/test_peer_gpu.py -m gpu`."""
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
    dp = _pkg("dp"); sp = _pkg("slidepath"); bpm = _pkg("bagpass"); synth = _pkg("synth"); peer_mod = _pkg("peer")
    res = dict(rank=rank)
    try:
        pg = peer_mod.PeerGroup(dev)
        # ---- 1) sharded inference: peer combine vs NCCL gather + combine vs the reference fixture
        case = load_case("mcat_concat_4096")
        cls = _pkg("mcat").MultimodalCoAttentionTransformer
        net = cls(omic_sizes=list(synth.OMIC_SIZES))
        net.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
        net = net.to(dev).eval()
        a, b = dp.patch_range(case["n"], rank, world)
        wsi = torch.from_numpy(case["bag"][a:b]).to(dev).to(torch.bfloat16)
        omics = [torch.from_numpy(o).to(dev) for o in case["omics"]]
        hz_n, _, _, amap_n = dp.sharded_inference(net, wsi, omics)
        hz_p, _, _, amap_p = dp.sharded_inference(net, wsi, omics, peer=pg)
        res["hz_nccl"], res["hz_peer"] = hz_n.cpu().numpy(), hz_p.cpu().numpy()
        res["amap_equal"] = bool(torch.equal(amap_n, amap_p))
        sh = dp.ShardedInference(net, wsi, omics, pg)
        for _ in range(3):
            hz_g, _, _, amap_g = sh.replay()
        torch.cuda.synchronize()
        res["hz_graph"] = hz_g.cpu().numpy()
        res["amap"] = amap_g.cpu().numpy()
        res["launches"] = sh.launches_per_replay
        # ---- 2) data-parallel optimizer step: peer reduce-scatter + Adam + all-gather vs NCCL all-reduce + flat Adam
        lens = [300, 517, 129, 1000]
        slides = [synth.make_slide(300 + i, n) for i, n in enumerate(lens)]
        mine = dp.slide_shard(len(slides), rank, world)
        pb = bpm.PackedBag.from_slides([torch.from_numpy(slides[i][0]).to(dev) for i in mine])
        om = [torch.stack([torch.from_numpy(slides[i][1][j]) for i in mine]).to(dev) for j in range(6)]
        labels = torch.tensor([slides[i][2] for i in mine], dtype=torch.int64, device=dev)
        cens = torch.tensor([slides[i][3] for i in mine], dtype=torch.float32, device=dev)
        finals = {}
        for mode in ("nccl", "peer"):
            net2 = cls(omic_sizes=list(synth.OMIC_SIZES))
            net2.load_state_dict({k: torch.from_numpy(v) for k, v in case["state"].items()})
            net2 = net2.to(dev).eval()
            tr = sp.BatchTrainer(net2, loss="nll", grad_acc_step=len(slides))
            tr.use_flat_adam(lr=2e-4, weight_decay=1e-5, peer=pg if mode == "peer" else None)
            off = tr.post_bucket_offset()
            for step in range(3):
                tr.step(pb, om, labels, cens, train=False)
                if mode == "nccl":
                    dist.all_reduce(tr.flat_grad)
                    tr.adam_step(zero_grad=True)
                else:                                   # two buckets, the step counter bumped by the second
                    tr.peer_adam_step(off, None, bump=False, slot=2)
                    tr.peer_adam_step(0, off, bump=True, slot=4)
            torch.cuda.synchronize()
            finals[mode] = tr.flat_param.clone()
            if mode == "peer":
                res["grad_zeroed"] = float(tr.flat_grad.abs().max().item())
        res["param_err"] = float((finals["peer"] - finals["nccl"]).norm() / finals["nccl"].norm())
        res["param_sum"] = float(finals["peer"].double().sum().item())
        pg.barrier()
        torch.cuda.synchronize()
        pg.close()
    except Exception as exc:      # report instead of leaving the parent waiting on the queue
        import traceback
        res["error"] = traceback.format_exc()
    q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_collectives_match_nccl_and_the_reference():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 90
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=300) for _ in range(world)], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
    for r in got:
        assert "error" not in r, r["error"]
    g = load_case("mcat_concat_4096")["gold"]
    for r in got:
        for k in ("hz_nccl", "hz_peer", "hz_graph"):
            assert np.max(np.abs(r[k] - g["hazards"]) / np.abs(g["hazards"])) < 1e-3, k
        assert np.allclose(r["hz_peer"], r["hz_nccl"], rtol=1e-6) and np.allclose(r["hz_graph"], r["hz_peer"], rtol=1e-6)
        assert r["amap_equal"] and r["launches"] > 0
        assert r["grad_zeroed"] == 0.0
        assert r["param_err"] < 1e-6, r["param_err"]
    amap = np.concatenate([r["amap"] for r in got], axis=1).astype(np.float64)
    ref = g["coattn"].astype(np.float64)
    assert np.max(np.abs(amap - ref) / (np.abs(ref) + 1e-3 * ref.max())) < 1e-3
    assert got[0]["param_sum"] == got[1]["param_sum"]          # bit-identical parameters on both ranks
