"""world_size-2 gloo test of the multi-GPU host logic (dp.py): gradient all-reduce over the flat buffer and the
all-gather + log-sum-exp combine of a patch-range-sharded bag (the combine arithmetic is checked here with the
oracle's math; the CUDA combine kernel is checked on the GPU in tests/test_multi_gpu.py)."""
import os
import sys
from importlib import import_module

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dp = import_module("multimodal-path-omic_b200.dp")
    # 1) gradient all-reduce: each rank contributes its slides' share, scaled by 1/global window
    n_slides, window = 10, 10
    mine = dp.slide_shard(n_slides, rank, world)
    g = torch.zeros(1000, dtype=torch.float32)
    for s in mine:
        g += torch.full((1000,), float(s + 1)) / window
    dp.all_reduce_gradients(g)
    # 2) sharded bag: per-rank softmax statistics -> all-gather
    rng = np.random.default_rng(0)
    N = 1000
    s_full = rng.standard_normal((6, N)) * 3
    h_full = rng.standard_normal((N, 256))
    a, b = dp.patch_range(N, rank, world)
    s_loc, h_loc = s_full[:, a:b], h_full[a:b]
    m = s_loc.max(axis=1, keepdims=True)
    lse = m[:, 0] + np.log(np.exp(s_loc - m).sum(axis=1))
    pooled = np.exp(s_loc - lse[:, None]) @ h_loc
    lse_all, pooled_all = dp.gather_shard_stats(torch.from_numpy(lse), torch.from_numpy(pooled))
    # 3) map export: every rank's [6, n_local] slice gathered on rank 0
    Nm = 1000
    full_map = torch.arange(6 * Nm, dtype=torch.float32).reshape(6, Nm)
    a, b = dp.patch_range(Nm, rank, world)
    gathered = dp.gather_attention_map(full_map[:, a:b].contiguous(), Nm)
    assert (gathered is None) == (rank != 0)
    if rank == 0:
        assert torch.equal(gathered, full_map)
    # ... and with an empty last rank (a 100-patch bag is one tile: rank 1 owns nothing)
    a, b = dp.patch_range(100, rank, world)
    small = dp.gather_attention_map(full_map[:, a:b].contiguous(), 100)
    if rank == 0:
        assert torch.equal(small, full_map[:, :100])
        out.put((g.numpy(), lse_all.numpy(), pooled_all.numpy(), s_full, h_full))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_gradient_allreduce_and_shard_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    g, lse_all, pooled_all, s_full, h_full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.allclose(g, sum(range(1, 11)) / 10.0)
    # combine (same arithmetic as mpo_lse_combine) must reproduce the unsharded softmax pooling
    M = lse_all.max(axis=0)
    w = np.exp(lse_all - M)
    lse = M + np.log(w.sum(axis=0))
    pooled = (pooled_all * w[:, :, None]).sum(axis=0) / w.sum(axis=0)[:, None]
    m = s_full.max(axis=1, keepdims=True)
    lse_ref = m[:, 0] + np.log(np.exp(s_full - m).sum(axis=1))
    pooled_ref = np.exp(s_full - lse_ref[:, None]) @ h_full
    assert np.allclose(lse, lse_ref, atol=1e-10) and np.allclose(pooled, pooled_ref, atol=1e-10)
