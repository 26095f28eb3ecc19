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


def sharded_inference(module, wsi_local, omics, group=None):
    """MCAT / NaCAGaT inference on one bag whose patches are split over the ranks of `group`.

    wsi_local: this rank's [n_local, 1024] slice (see patch_range; ranks at the tail of a short bag may hold ZERO
    rows); omics: the same 6 vectors on every rank.  Returns hazards, S, Y (identical on every rank) and this rank's
    [6, n_local] slice of the co-attention map.  Every rank takes part in the all-gather: an empty rank contributes
    lse = -inf / pooled = 0, which the log-sum-exp combine weights with exactly zero."""
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
