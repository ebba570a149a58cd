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


def gather_trace_columns(local_values, n_total, dst=0, group=None):
    """Collect the per-proof trace columns `[n_local, n_cols, n_rows]` (int32/uint32 words) of every rank's block on rank `dst`
    (dst=None: on every rank), in proof order -- for a caller that wants one device to hold the whole batch's traces (north_star:
    "NCCL ... to gather per-proof verdicts and trace columns").  Blocks differ by at most one proof, so the exchange is one
    all-gather / gather of equally sized, padded buffers: no staging through the host, NVLink / NVSwitch carries it under NCCL.
    Returns the `[n_total, n_cols, n_rows]` tensor (None on the ranks that are not `dst`)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_values
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [hi - lo for lo, hi in (shard_range(n_total, r, world) for r in range(world))]
    if local_values.shape[0] != sizes[rank]:
        raise ValueError("local block has %d proofs, expected %d" % (local_values.shape[0], sizes[rank]))
    width = max(sizes)
    send = local_values
    if sizes[rank] < width:                                  # pad the short blocks by one proof
        send = torch.zeros((width,) + tuple(local_values.shape[1:]), dtype=local_values.dtype, device=local_values.device)
        send[: sizes[rank]] = local_values
    send = send.contiguous()
    if dst is None:
        out = torch.empty((world * width,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(out, send, group=group)
    else:
        out = torch.empty((world * width,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device) if rank == dst else None
        dist.gather(send, list(out.split(width)) if rank == dst else None, dst=dst, group=group)
        if rank != dst:
            return None
    if all(s == width for s in sizes):
        return out
    return torch.cat([out[r * width: r * width + sizes[r]] for r in range(world)])
