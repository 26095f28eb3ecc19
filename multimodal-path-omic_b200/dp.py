"""Multi-GPU drivers, one process per GPU over torch.distributed (NCCL on the B200 box, gloo in CPU tests).

1. Data-parallel training over slides: rank r owns slides r, r+G, ...; every rank accumulates its
   share of the gradient window locally (BatchTrainer) and ONE all-reduce of the flat fp32 gradient buffer follows
   per optimizer step.  The reference has no distributed code (its nn.DataParallel wrapper is a no-op at batch 1,
   models/mcat/main.py:267-268), so the loop semantics reproduced are those of its single-GPU driver
   (main.py:30-74): loss / grad_acc_step, optimizer step once per window.
2. Patch-range sharding of one very large bag: rank r streams patches [r*N/G, (r+1)*N/G), the
   per-query (lse, pooled) of all ranks are all-gathered (6 x 257 floats each) and merged by mpo_lse_combine on
   every rank; the tail is replicated; each rank owns its slice of the [6, N] attention map.
"""
import torch
import torch.distributed as dist

TILE = 128


def slide_shard(num_slides, rank, world):
    """indices of the slides rank `rank` processes (round-robin, like a DistributedSampler without padding)."""
    return list(range(rank, num_slides, world))


def patch_range(num_patches, rank, world):
    """[start, end) patch rows of rank `rank`: contiguous, tile-aligned starts, ranks at the tail may be empty."""
    tiles = (num_patches + TILE - 1) // TILE
    per = (tiles + world - 1) // world * TILE
    start = min(num_patches, rank * per)
    end = min(num_patches, (rank + 1) * per)
    return start, end


def all_reduce_gradients(flat_grad, group=None):
    """sum the flat gradient buffer over ranks (the loss is already scaled by 1/global window)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def gather_shard_stats(lse_local, pooled_local, group=None):
    """all-gather each rank's (lse [6], pooled [6,256]) -> ([G,6], [G,6,256]) on every rank."""
    world = dist.get_world_size(group)
    packed = torch.cat([lse_local.reshape(6, 1), pooled_local.reshape(6, -1)], dim=1).contiguous()   # [6, 257]
    out = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(out, packed, group=group)
    allp = torch.stack(out)                       # [G, 6, 257]
    return allp[:, :, 0].contiguous(), allp[:, :, 1:].contiguous()


class DataParallelTrainer:
    """BatchTrainer + one gradient all-reduce per optimizer step."""

    def __init__(self, module, optimizer, slides_per_rank_per_step, loss="nll", group=None):
        from . import slidepath
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.trainer = slidepath.BatchTrainer(module, loss=loss,
                                              grad_acc_step=slides_per_rank_per_step * self.world)
        self.optimizer = optimizer

    def step(self, bag, omics, labels, censor, train=True):
        loss, hazards, S = self.trainer.step(bag, omics, labels, censor, train=train)
        all_reduce_gradients(self.trainer.flat_grad, self.group)
        self.optimizer.step()
        self.trainer.zero_grad()
        return loss, hazards, S


def sharded_inference(module, wsi_local, omics, group=None, peer=None):
    """MCAT / NaCAGaT inference on one bag whose patches are split over the ranks of `group`.

    wsi_local: this rank's [n_local, 1024] slice (see patch_range; ranks at the tail of a short bag may hold ZERO
    rows); omics: the same 6 vectors on every rank.  Returns hazards, S, Y (identical on every rank) and this rank's
    [6, n_local] slice of the co-attention map.  Every rank takes part in the all-gather: an empty rank contributes
    lse = -inf / pooled = 0, which the log-sum-exp combine weights with exactly zero.
    peer (peer.PeerGroup): the exchange + merge run as ONE kernel over NVLink peer memory (mpo_peer_lse_combine) instead of
    an NCCL all-gather followed by mpo_lse_combine."""
    from . import bagpass as bp
    eng = module._engine
    n_local = int(wsi_local.shape[-2])
    if n_local == 0:
        # the kernels need at least one row to launch on: one zero row whose statistics are discarded below
        wsi_run = torch.zeros((1, wsi_local.shape[-1]), dtype=torch.bfloat16, device=wsi_local.device)
    else:
        wsi_run = wsi_local
    bag = bp.PackedBag.from_slides([wsi_run])
    model = eng.binding.build(grads=None)

    def combine(st):
        lse_l, pooled_l = st.bag_ws.lse[0], st.bag_ws.pooled[0]
        if n_local == 0:
            lse_l = torch.full_like(lse_l, float("-inf"))
            pooled_l = torch.zeros_like(pooled_l)
        if peer is not None:
            peer.lse_combine(lse_l.contiguous(), pooled_l.contiguous(), st.bag_ws.lse[0], st.bag_ws.pooled[0])
            return
        lse_all, pooled_all = gather_shard_stats(lse_l, pooled_l, group)
        lse, pooled = bp.lse_combine(lse_all, pooled_all)
        st.bag_ws.lse.copy_(lse.reshape(1, 6))
        st.bag_ws.pooled.copy_(pooled.reshape(1, 6, -1))

    with torch.no_grad():
        st = eng.forward(model, bag, [o.reshape(1, -1) for o in omics], train=False, save_for_backward=False,
                         after_bag=combine)
        amap = eng.attention_map(st)
    if n_local == 0:
        amap = amap[:, :0]
    return st.hazards, st.S, st.Y, amap


class ShardedInference:
    """A patch-range sharded inference call captured as ONE CUDA graph per rank (BASELINE config 5): SNN / query fold ->
    bag forward over this rank's patches -> mpo_peer_lse_combine over NVLink -> replicated tail -> this rank's slice of
    the co-attention map.  The caller refreshes `wsi_local` / `omics` in place and calls replay().

    wsi_local: static bf16 [n_local, 1024] device tensor (n_local may be 0); omics: 6 static fp32 device vectors."""

    def __init__(self, module, wsi_local, omics, peer):
        from . import bagpass as bp
        self.eng = module._engine
        self.peer = peer
        self.n_local = int(wsi_local.shape[0])
        dev = wsi_local.device
        if self.n_local == 0:
            wsi_local = torch.zeros((1, wsi_local.shape[-1]), dtype=torch.bfloat16, device=dev)
        if wsi_local.dtype != torch.bfloat16:
            raise RuntimeError("ShardedInference streams the shard where it lies: pass a bf16 tensor")
        self.bag = bp.PackedBag.from_slides([wsi_local])
        self.omics = [o.reshape(1, -1) for o in omics]
        self.model = self.eng.binding.build(grads=None)
        self.st = self.eng.alloc_state(self.model, self.bag, save_for_backward=False)
        self.amap = torch.empty((6, self.bag.total_rows), dtype=torch.float32, device=dev)
        self._neg_inf = torch.full((6,), float("-inf"), dtype=torch.float32, device=dev)
        self._zeros = torch.zeros((6, 256), dtype=torch.float32, device=dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                       # warm-up outside the capture (same call count on every rank)
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        _lib_mod = __import__("importlib").import_module(__package__ + "._lib")
        _lib_mod.lib().mpo_launch_count(1)
        with torch.cuda.graph(self.graph), torch.no_grad():
            self._run()
        self.launches_per_replay = int(_lib_mod.lib().mpo_launch_count(1))

    def _combine(self, st):
        lse_l, pooled_l = st.bag_ws.lse[0], st.bag_ws.pooled[0]
        if self.n_local == 0:
            lse_l, pooled_l = self._neg_inf, self._zeros
        self.peer.lse_combine(lse_l, pooled_l, st.bag_ws.lse[0], st.bag_ws.pooled[0])

    def _run(self):
        from . import bagpass as bp
        self.eng.forward(self.model, self.bag, self.omics, train=False, save_for_backward=False, st=self.st,
                         after_bag=self._combine)
        bp.attention_map(self.bag, self.st.bag_ws, out=self.amap)

    def replay(self):
        """-> hazards, S, Y [1, K] (identical on every rank) and this rank's [6, n_local] slice of the map."""
        self.graph.replay()
        st = self.st
        return st.hazards, st.S, st.Y, (self.amap if self.n_local else self.amap[:, :0])


def gather_attention_map(amap_local, num_patches, group=None, dst=0):
    """Patch-range sharded inference leaves every rank with its [6, n_local] slice of the co-attention map; this gathers
    the full [6, num_patches] map on rank `dst` (4.8 MB at 200 000 patches), None elsewhere.
    Slices follow patch_range(): rank r owns columns [r * per, (r + 1) * per)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = patch_range(num_patches, 0, world)[1] if world > 0 else num_patches
    per = max(per, 1)
    a, b = patch_range(num_patches, rank, world)
    pad = torch.zeros((amap_local.shape[0], per), dtype=amap_local.dtype, device=amap_local.device)
    pad[:, :b - a] = amap_local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst, group=group)
    if rank != dst:
        return None
    full = torch.cat(parts, dim=1)[:, :num_patches]
    return full.contiguous()
