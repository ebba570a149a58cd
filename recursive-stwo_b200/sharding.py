"""Multi-GPU partitioning of a proof batch: proofs are independent units, so each rank verifies a
contiguous block and the only collective is the gather of the per-proof verdict bytes (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  The reference has no equivalent (single-threaded, examples/multi-proofs/src/main.rs:67-139
shares one constraint system); this is the layer north_star adds."""
import ctypes

import numpy as np

from . import _lib


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank` out of `world`; blocks differ by at most one item (stwo_b200_shard_range)."""
    lo, hi = ctypes.c_uint64(0), ctypes.c_uint64(0)
    if not 0 <= rank < world or _lib.load().stwo_b200_shard_range(n_items, rank, world, ctypes.byref(lo), ctypes.byref(hi)) != _lib.OK:
        raise ValueError("rank outside the world")
    return int(lo.value), int(hi.value)


def shard_by_work(work, world):
    """Contiguous blocks of a list of units with unequal cost (proofs of different shapes: cost ~ permutations x rows): cut points at
    the multiples of total / world of the running cost, so no rank is pinned by the expensive units.  -> [(lo, hi)] * world"""
    work = np.asarray(work, dtype=np.float64)
    cum = np.concatenate([[0.0], np.cumsum(work)])
    cuts = [int(np.searchsorted(cum, cum[-1] * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, len(work)
    for r in range(1, world + 1):
        cuts[r] = max(cuts[r], cuts[r - 1])
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class Comm:
    """NCCL communicator created through the library's own C entry points (stwo_b200_comm_*): what a non-Python host uses.  The
    128-byte unique id travels over any side channel; here torch.distributed (already initialised by the launcher) broadcasts it."""

    def __init__(self, rank, world, device):
        import torch
        import torch.distributed as dist
        self.rank, self.world = rank, world
        idbuf = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _lib.call("stwo_b200_comm_unique_id", idbuf.ctypes.data_as(ctypes.c_void_p))
        t = torch.from_numpy(idbuf).to(device)
        if world > 1:
            dist.broadcast(t, src=0)
        idbuf = t.cpu().numpy()
        self._h = ctypes.c_void_p()
        _lib.call("stwo_b200_comm_init", idbuf.ctypes.data_as(ctypes.c_void_p), rank, world, ctypes.byref(self._h))

    def gather_verdicts(self, local_verdict, local_stage, n_total):
        import torch
        from .hashing import _dptr, _stream
        dev = local_verdict.device
        verdict = torch.empty(n_total, dtype=torch.uint8, device=dev)
        stage = torch.empty(n_total, dtype=torch.uint8, device=dev)
        nbytes = int(_lib.load().stwo_b200_gather_verdicts_scratch_bytes(self.world, n_total))
        scratch = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        _lib.call("stwo_b200_gather_verdicts", self._h, self.rank, self.world, n_total, _dptr(local_verdict), _dptr(local_stage), _dptr(verdict),
                  _dptr(stage), _dptr(scratch), nbytes, _stream())
        return verdict, stage

    def gather_trace_columns(self, local_values, n_total, dst=0):
        import torch
        from .hashing import _dptr, _stream
        words = int(np.prod(local_values.shape[1:]))
        out = torch.empty((n_total,) + tuple(local_values.shape[1:]), dtype=local_values.dtype, device=local_values.device) if self.rank == dst else None
        _lib.call("stwo_b200_gather_trace_columns", self._h, self.rank, self.world, dst, n_total, words, _dptr(local_values.contiguous()),
                  _dptr(out), _stream())
        return out

    def close(self):
        if self._h:
            _lib.load().stwo_b200_comm_destroy(self._h)
            self._h = None


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
