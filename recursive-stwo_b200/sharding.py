"""Multi-GPU partitioning of a proof batch: proofs are independent units, so each rank verifies a
contiguous block and the only collective is the gather of the per-proof verdict bytes (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The reference has no equivalent (single-threaded, examples/multi-proofs/src/main.rs:67-139
shares one constraint system); this is the layer north_star adds."""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank` out of `world`; blocks differ by at most one item."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_verdicts(local_verdict, local_stage, n_total, group=None):
    """All-gather the (verdict, stage) bytes of every rank's block into full-length tensors on every rank.
    local_*: uint8 tensors (CUDA under NCCL, CPU under gloo) of this rank's block."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_verdict, local_stage
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.full((2, width), 255, dtype=torch.uint8, device=local_verdict.device)
    lo, hi = sizes[rank]
    if local_verdict.numel() != hi - lo:
        raise ValueError("local block has %d items, expected %d" % (local_verdict.numel(), hi - lo))
    pad[0, : hi - lo] = local_verdict
    pad[1, : hi - lo] = local_stage
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    verdict = torch.cat([out[r][0, : sizes[r][1] - sizes[r][0]] for r in range(world)])
    stage = torch.cat([out[r][1, : sizes[r][1] - sizes[r][0]] for r in range(world)])
    return verdict, stage
